"""Entry-point mirrors of the reference's two inference scripts.

  inference_all_seg  <- 分割/predict.py:11-51  (UNet over a loader; masks returned / optionally painted)
  inference_all_cls  <- 分类/test.py:74-96     (GoogLeNet over a loader of ROI crops; sorted result.txt)

Same arguments and on-disk outputs as the reference functions for the parts on the hot path; the reference's
Excel dump (predict.py:50-51) is host I/O outside the path and is not reproduced."""
import os

import numpy as np
import torch


@torch.no_grad()
def inference_all_cls(model, test_loader, device, save_dir="test_results"):
    model.eval()
    os.makedirs(save_dir, exist_ok=True)
    records = []
    for data in test_loader:
        imgs = data["image"].float().to(device)
        cl_out = model(imgs)
        pred = torch.argmax(torch.softmax(cl_out, dim=1), dim=1).cpu().numpy()
        for i in range(imgs.size(0)):
            records.append(f"{data['filename'][i].replace('.png', '')} {int(pred[i])}")
    records.sort(key=lambda x: int(x.split()[0].replace(".jpg", "").replace(".png", "")))
    with open(os.path.join(save_dir, "result.txt"), "w") as f:
        for line in records:
            f.write(line + "\n")
    return records


@torch.no_grad()
def inference_all_seg(model, test_loader, device, save_dir=None):
    """Returns {filename: uint8 mask [224,224]}; with save_dir, also writes what predict.py:13-45 writes: one RGB PNG
    per image under `save_dir/Segmentation_Results/`, named `filename.replace('.jpg', '') + '.png'`, red (255,0,0) where
    mask == 1 on black (the reference's 50 176-iteration putpixel loop as one vectorised NumPy write)."""
    model.eval()
    out = {}
    seg_dir = None
    if save_dir is not None:
        seg_dir = os.path.join(save_dir, "Segmentation_Results")
        os.makedirs(seg_dir, exist_ok=True)
    for data in test_loader:
        imgs = data["image"].float().to(device)
        _, masks, _ = model.forward_mask_boxes(imgs)
        masks = masks.cpu().numpy()
        for i, filename in enumerate(data["filename"]):
            out[filename] = masks[i]
            if seg_dir is not None:
                from PIL import Image
                canvas = np.zeros(masks[i].shape + (3,), np.uint8)
                canvas[masks[i] == 1] = (255, 0, 0)
                Image.fromarray(canvas, "RGB").save(os.path.join(seg_dir, filename.replace(".jpg", "") + ".png"))
    return out


@torch.no_grad()
def grade_images(pipeline, gray_images, filenames, save_dir=None, batch_size=256):
    """The whole of 分类/test.py:122-134 + :74-96 on the device for a list of grayscale images of one size:
    cv2.imread(path, 0) output (uint8 [H, W]) -> wavelet_enhance -> resize 224 + to_tensor -> UNet -> mask -> bbox ->
    ROI crop/resize -> GoogLeNet -> argmax; returns the sorted "<name> <class>" records (and writes result.txt).

    pipeline: pipeline.TwoStagePipeline (or lower.PipelineRunner); gray_images: uint8 array / tensor [N, H, W]."""
    from .util.wavelet import wavelet_enhance_batch
    runner = getattr(pipeline, "runner", pipeline)
    g = torch.as_tensor(np.ascontiguousarray(gray_images) if isinstance(gray_images, np.ndarray) else gray_images)
    records = []
    for s in range(0, g.shape[0], batch_size):
        rgb = wavelet_enhance_batch(g[s:s + batch_size].to(runner.dev))      # [B,H,W,3] uint8, test.py:128-129
        _, _, cls = runner(rgb)                                              # test.py:130-131 and :82-84
        pred = torch.argmax(torch.softmax(cls, dim=1), dim=1).cpu().numpy()  # test.py:86
        for i, name in enumerate(filenames[s:s + batch_size]):
            records.append(f"{name.replace('.png', '')} {int(pred[i])}")
    records.sort(key=lambda x: int(x.split()[0].replace(".jpg", "").replace(".png", "")))
    if save_dir is not None:
        os.makedirs(save_dir, exist_ok=True)
        with open(os.path.join(save_dir, "result.txt"), "w") as f:
            for line in records:
                f.write(line + "\n")
    return records


inference_all = inference_all_cls
