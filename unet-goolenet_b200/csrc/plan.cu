// Plan images: a compiled program (op list), its device memory layout and its constant data (packed weights) as ONE
// relocatable byte image, so that a host WITHOUT the Python lowering can run the path through the C ABI alone
// (ug_plan_load / ug_plan_copy_in / ug_plan_run / ug_plan_copy_out; include/ugnet.h).  The image is produced once by
// the Python tooling (unet-goolenet_b200/engine.py export_plan, lower.PipelineRunner.export_plan).
//
// Layout (little endian, all tables 8-byte aligned):
//   header  { char magic[8] = "UGPLAN01"; u32 n_allocs, n_ops, n_relocs, n_io; u64 op_bytes (= sizeof(ug_op)) }
//   allocs  n_allocs x { u64 bytes; u64 init_offset (into the image, 0 = uninitialised); u64 init_bytes }
//   ops     n_ops x ug_op (pointer fields hold stale addresses of the exporting process)
//   relocs  n_relocs x { u32 op; u32 field_offset (bytes into the ug_op); u32 alloc; u32 pad; u64 offset }
//   io      n_io x { char name[32]; u32 alloc; u32 pad; u64 offset; u64 bytes }
//   blobs   initial contents
#include <cstring>
#include <new>
#include <string>
#include <vector>
#include "engine.h"

namespace {

struct PlanHeader {
  char magic[8];
  uint32_t n_allocs, n_ops, n_relocs, n_io;
  uint64_t op_bytes;
};
struct PlanAlloc {
  uint64_t bytes, init_offset, init_bytes;
};
struct PlanReloc {
  uint32_t op, field_offset, alloc, pad;
  uint64_t offset;
};
struct PlanIo {
  char name[32];
  uint32_t alloc, pad;
  uint64_t offset, bytes;
};

}  // namespace

struct ug_plan_s {
  void* arena = nullptr;
  size_t arena_bytes = 0;
  ug_program program = nullptr;
  struct Io {
    std::string name;
    void* ptr;
    size_t bytes;
  };
  std::vector<Io> io;
};

using namespace ug;

extern "C" {

int ug_plan_load(ug_handle h, const void* image, size_t image_bytes, ug_plan* out) {
  if (!h || !image || !out) return UG_EINVAL;
  *out = nullptr;
  DeviceGuard guard(h);
  const unsigned char* base = static_cast<const unsigned char*>(image);
  if (image_bytes < sizeof(PlanHeader)) return set_error(h, UG_EINVAL, "plan: image too small");
  PlanHeader hd;
  memcpy(&hd, base, sizeof(hd));
  if (memcmp(hd.magic, "UGPLAN01", 8) != 0) return set_error(h, UG_EINVAL, "plan: bad magic");
  if (hd.op_bytes != sizeof(ug_op))
    return set_error(h, UG_EINVAL, "plan: image built for sizeof(ug_op) = %llu, this library has %zu",
                     (unsigned long long)hd.op_bytes, sizeof(ug_op));
  size_t off = sizeof(PlanHeader);
  const size_t need = off + (size_t)hd.n_allocs * sizeof(PlanAlloc) + (size_t)hd.n_ops * sizeof(ug_op) +
                      (size_t)hd.n_relocs * sizeof(PlanReloc) + (size_t)hd.n_io * sizeof(PlanIo);
  if (hd.n_ops == 0 || hd.n_allocs == 0 || need > image_bytes) return set_error(h, UG_EINVAL, "plan: truncated tables");
  std::vector<PlanAlloc> allocs(hd.n_allocs);
  memcpy(allocs.data(), base + off, allocs.size() * sizeof(PlanAlloc));
  off += allocs.size() * sizeof(PlanAlloc);
  std::vector<ug_op> ops(hd.n_ops);
  memcpy(ops.data(), base + off, ops.size() * sizeof(ug_op));
  off += ops.size() * sizeof(ug_op);
  std::vector<PlanReloc> relocs(hd.n_relocs);
  if (hd.n_relocs) memcpy(relocs.data(), base + off, relocs.size() * sizeof(PlanReloc));
  off += relocs.size() * sizeof(PlanReloc);
  std::vector<PlanIo> ios(hd.n_io);
  if (hd.n_io) memcpy(ios.data(), base + off, ios.size() * sizeof(PlanIo));

  // one arena, every allocation 256-byte aligned
  std::vector<size_t> start(hd.n_allocs);
  size_t total = 0;
  for (uint32_t i = 0; i < hd.n_allocs; ++i) {
    total = (total + 255) & ~(size_t)255;
    start[i] = total;
    total += allocs[i].bytes;
    if (allocs[i].init_bytes > allocs[i].bytes || (allocs[i].init_bytes && allocs[i].init_offset + allocs[i].init_bytes > image_bytes))
      return set_error(h, UG_EINVAL, "plan: allocation %u has an out-of-range initialiser", i);
  }
  ug_plan_s* p = new (std::nothrow) ug_plan_s();
  if (!p) return UG_ENOMEM;
  cudaError_t ce = cudaMalloc(&p->arena, total ? total : 256);
  if (ce != cudaSuccess) {
    delete p;
    return set_error(h, UG_ENOMEM, "plan: cudaMalloc of %zu bytes failed: %s", total, cudaGetErrorString(ce));
  }
  p->arena_bytes = total;
  char* arena = static_cast<char*>(p->arena);
  int rc = UG_OK;
  for (uint32_t i = 0; i < hd.n_allocs && rc == UG_OK; ++i)
    if (allocs[i].init_bytes)
      rc = check_cuda(h, cudaMemcpy(arena + start[i], base + allocs[i].init_offset, allocs[i].init_bytes, cudaMemcpyHostToDevice),
                      "plan: constant upload");
  for (uint32_t i = 0; i < hd.n_relocs && rc == UG_OK; ++i) {
    const PlanReloc& r = relocs[i];
    if (r.op >= hd.n_ops || r.alloc >= hd.n_allocs || r.field_offset + sizeof(void*) > sizeof(ug_op) || r.offset > allocs[r.alloc].bytes) {
      rc = set_error(h, UG_EINVAL, "plan: relocation %u out of range", i);
      break;
    }
    void* v = arena + start[r.alloc] + r.offset;
    memcpy(reinterpret_cast<char*>(&ops[r.op]) + r.field_offset, &v, sizeof(void*));
  }
  for (uint32_t i = 0; i < hd.n_io && rc == UG_OK; ++i) {
    if (ios[i].alloc >= hd.n_allocs || ios[i].offset + ios[i].bytes > allocs[ios[i].alloc].bytes) {
      rc = set_error(h, UG_EINVAL, "plan: io entry %u out of range", i);
      break;
    }
    ios[i].name[31] = 0;
    p->io.push_back({ios[i].name, arena + start[ios[i].alloc] + ios[i].offset, (size_t)ios[i].bytes});
  }
  if (rc == UG_OK) rc = ug_program_create(h, ops.data(), (int)ops.size(), &p->program);
  if (rc != UG_OK) {
    cudaFree(p->arena);
    delete p;
    return rc;
  }
  *out = p;
  return UG_OK;
}

int ug_plan_io(ug_plan p, const char* name, void** dev_ptr, size_t* bytes) {
  if (!p || !name) return UG_EINVAL;
  for (const auto& e : p->io)
    if (e.name == name) {
      if (dev_ptr) *dev_ptr = e.ptr;
      if (bytes) *bytes = e.bytes;
      return UG_OK;
    }
  return UG_EINVAL;
}

int ug_plan_num_io(ug_plan p) { return p ? (int)p->io.size() : 0; }

const char* ug_plan_io_name(ug_plan p, int i) { return (p && i >= 0 && i < (int)p->io.size()) ? p->io[i].name.c_str() : nullptr; }

ug_program ug_plan_program(ug_plan p) { return p ? p->program : nullptr; }

size_t ug_plan_device_bytes(ug_plan p) { return p ? p->arena_bytes : 0; }

int ug_plan_copy_in(ug_handle h, ug_plan p, const char* name, const void* host, size_t bytes, void* stream) {
  if (!h || !p || !host) return UG_EINVAL;
  DeviceGuard guard(h);
  void* dst = nullptr;
  size_t cap = 0;
  if (ug_plan_io(p, name, &dst, &cap) != UG_OK) return set_error(h, UG_EINVAL, "plan: no io buffer named '%s'", name ? name : "");
  if (bytes > cap) return set_error(h, UG_EINVAL, "plan: '%s' holds %zu bytes, %zu given", name, cap, bytes);
  return check_cuda(h, cudaMemcpyAsync(dst, host, bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)), "plan: copy in");
}

int ug_plan_copy_out(ug_handle h, ug_plan p, const char* name, void* host, size_t bytes, void* stream) {
  if (!h || !p || !host) return UG_EINVAL;
  DeviceGuard guard(h);
  void* src = nullptr;
  size_t cap = 0;
  if (ug_plan_io(p, name, &src, &cap) != UG_OK) return set_error(h, UG_EINVAL, "plan: no io buffer named '%s'", name ? name : "");
  if (bytes > cap) return set_error(h, UG_EINVAL, "plan: '%s' holds %zu bytes, %zu requested", name, cap, bytes);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = check_cuda(h, cudaMemcpyAsync(host, src, bytes, cudaMemcpyDeviceToHost, s), "plan: copy out");
  if (rc == UG_OK) rc = check_cuda(h, cudaStreamSynchronize(s), "plan: copy out sync");
  return rc;
}

int ug_plan_run(ug_handle h, ug_plan p, void* stream) {
  if (!h || !p) return UG_EINVAL;
  return ug_program_run(h, p->program, stream);
}

int ug_plan_destroy(ug_handle h, ug_plan p) {
  if (!p) return UG_OK;
  DeviceGuard guard(h);
  if (p->program) ug_program_destroy(h, p->program);
  if (p->arena) cudaFree(p->arena);
  delete p;
  return UG_OK;
}

}  // extern "C"
