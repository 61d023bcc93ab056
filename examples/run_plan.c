/* run_plan.c — the two-stage path from a host that has neither Python nor the CUDA runtime headers: plain C against
 * include/ugnet.h and libugnet.so.
 *
 *   gcc -O2 -I include examples/run_plan.c -o run_plan -L unet-goolenet_b200 -lugnet -Wl,-rpath,$PWD/unet-goolenet_b200
 *   ./run_plan plan.bin images.f32 B out_prefix
 *
 * plan.bin    a plan image exported once by the Python tooling: PipelineRunner(unet_sd, googlenet_sd, dev).export_plan(B)
 * images.f32  B x 3 x 224 x 224 float32 in [0,1] (what 分类/test.py:130 feeds process_and_augment_roi)
 * writes      <out_prefix>.mask.u8 (B x 224 x 224), <out_prefix>.boxes.i32 (B x 4: x0 y0 x1 y1), <out_prefix>.cls.f32 (B x 6)
 * i.e. the results of UNetTaskAligWeight.forward -> roi.py:22-36 -> roi.py:39-49 -> GoogLeNetClassifier.forward. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ugnet.h"

static void* read_file(const char* path, size_t* n) {
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
  fseek(f, 0, SEEK_END);
  *n = (size_t)ftell(f);
  fseek(f, 0, SEEK_SET);
  void* p = malloc(*n);
  if (!p || fread(p, 1, *n, f) != *n) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
  fclose(f);
  return p;
}

static void write_file(const char* prefix, const char* suffix, const void* p, size_t n) {
  char path[1024];
  snprintf(path, sizeof(path), "%s.%s", prefix, suffix);
  FILE* f = fopen(path, "wb");
  if (!f || fwrite(p, 1, n, f) != n) { fprintf(stderr, "cannot write %s\n", path); exit(2); }
  fclose(f);
}

#define CHECK(call)                                                                         \
  do {                                                                                      \
    int rc_ = (call);                                                                       \
    if (rc_ != UG_OK) {                                                                     \
      fprintf(stderr, "%s failed: %d (%s)\n", #call, rc_, h ? ug_last_error(h) : "no handle"); \
      return 1;                                                                             \
    }                                                                                       \
  } while (0)

int main(int argc, char** argv) {
  if (argc != 5) { fprintf(stderr, "usage: %s plan.bin images.f32 B out_prefix\n", argv[0]); return 2; }
  const int B = atoi(argv[3]);
  size_t plan_bytes = 0, img_bytes = 0;
  void* image = read_file(argv[1], &plan_bytes);
  float* imgs = (float*)read_file(argv[2], &img_bytes);
  if (img_bytes != (size_t)B * 3 * 224 * 224 * sizeof(float)) { fprintf(stderr, "images.f32 has the wrong size\n"); return 2; }

  ug_handle h = NULL;
  ug_plan plan = NULL;
  CHECK(ug_create(0, &h));
  CHECK(ug_plan_load(h, image, plan_bytes, &plan));
  free(image);
  printf("plan: %d named buffers, %.1f MB of device memory\n", ug_plan_num_io(plan), ug_plan_device_bytes(plan) / 1e6);

  unsigned char* mask = (unsigned char*)malloc((size_t)B * 224 * 224);
  int* boxes = (int*)malloc((size_t)B * 4 * sizeof(int));
  float* cls = (float*)malloc((size_t)B * 6 * sizeof(float));
  CHECK(ug_plan_copy_in(h, plan, "x_in", imgs, img_bytes, NULL));
  CHECK(ug_plan_run(h, plan, NULL));
  CHECK(ug_plan_copy_out(h, plan, "mask", mask, (size_t)B * 224 * 224, NULL));
  CHECK(ug_plan_copy_out(h, plan, "boxes", boxes, (size_t)B * 4 * sizeof(int), NULL));
  CHECK(ug_plan_copy_out(h, plan, "cls_logits", cls, (size_t)B * 6 * sizeof(float), NULL));
  write_file(argv[4], "mask.u8", mask, (size_t)B * 224 * 224);
  write_file(argv[4], "boxes.i32", boxes, (size_t)B * 4 * sizeof(int));
  write_file(argv[4], "cls.f32", cls, (size_t)B * 6 * sizeof(float));
  for (int i = 0; i < B; ++i) {
    int best = 0;
    for (int k = 1; k < 6; ++k)
      if (cls[i * 6 + k] > cls[i * 6 + best]) best = k;   /* argmax(softmax(logits)), 分类/test.py:86 */
    printf("image %d: box (%d,%d)-(%d,%d) class %d\n", i, boxes[4 * i], boxes[4 * i + 1], boxes[4 * i + 2], boxes[4 * i + 3], best);
  }
  CHECK(ug_plan_destroy(h, plan));
  ug_destroy(h);
  free(imgs); free(mask); free(boxes); free(cls);
  return 0;
}
