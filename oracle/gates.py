"""ORACLE-side parity gates (test infrastructure only — used by tests/, __graft_entry__.smoke() and bench.py's
in-run correctness check; never imported by the product package).

`pipeline_gates` runs the fp32 restatement of the reference path (oracle/unet_ref.py, roi_ref.py, googlenet_ref.py)
on the same images and weights as the engine and evaluates BASELINE.json's contract:

  * segmentation masks >= 99.9 % pixel agreement with the reference (分割/nets/basicUnet.py:406-437 + roi.py:22-23);
  * bbox / crop bit-exact given the same mask: the engine's box must equal the reference rule (roi.py:25-36) applied to
    the engine's own mask, and the engine's uint8 crop must equal the reference crop/resize chain (roi.py:39-44,
    data_utils.py:102,146-147) for that box — for every image; where the engine's mask equals the reference mask this
    is the reference's box and crop;
  * class logits within 1e-2 of the per-image logit scale with identical argmax (on every image whose reference top-2
    margin is outside the 2 x 1e-2 tolerance band — see argmax_gate), the reference classifier
    (分类/test.py:64-73) being evaluated on the reference crop of the engine's box (the crop depends on the image and
    the box only, so for every image whose box equals the reference's this is the unmodified reference path).

The network oracles are plain PyTorch and run on whatever device the state_dict lives on: on the GPU box they are run
in fp32 on the GPU (TF32 off) so that BASELINE-size batches take seconds; the integer ROI oracle is NumPy on the host.
"""
import numpy as np
import torch

from . import googlenet_ref, roi_ref, unet_ref

MASK_AGREEMENT = 0.999      # BASELINE.json north_star: >= 99.9 % pixel agreement
LOGIT_REL = 1e-2            # logits within 1e-2 relative (to the per-image logit scale), identical argmax


class _Fp32:
    """TF32 off for the duration of an oracle evaluation on the GPU."""

    def __enter__(self):
        self.prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False

    def __exit__(self, *a):
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = self.prev


def oracle_unet_logits(unet_sd, imgs, device, chunk=16):
    """fp32 reference logits [B,1,224,224] (CPU tensor) of float images [B,3,224,224] (NumPy or tensor)."""
    sd = {k: v.to(device) for k, v in unet_sd.items()}
    x = torch.as_tensor(imgs)
    out = []
    with torch.no_grad(), _Fp32():
        for s in range(0, x.shape[0], chunk):
            out.append(unet_ref.unet_forward(sd, x[s:s + chunk].to(device)).float().cpu())
    return torch.cat(out)


def oracle_googlenet_logits(gnet_sd, crops, device, chunk=64):
    sd = {k: v.to(device) for k, v in gnet_sd.items()}
    x = torch.as_tensor(crops)
    out = []
    with torch.no_grad(), _Fp32():
        for s in range(0, x.shape[0], chunk):
            out.append(googlenet_ref.googlenet_forward(sd, x[s:s + chunk].to(device)).float().cpu())
    return torch.cat(out)


def logit_rel_err(got, ref):
    """max|d| / max|ref| per image."""
    return ((got - ref).abs().amax(1) / ref.abs().amax(1)).numpy()


def argmax_gate(got, ref):
    """-> (number of images with identical argmax, number of DECIDED images, all decided images identical).
    An image is decided when the reference's top-2 margin exceeds twice the logit tolerance (2 * 1e-2 * scale): inside
    that band two logits that each moved by <= 1e-2 * scale may legitimately swap, outside it a different argmax is an
    error whatever the logit error."""
    top2 = ref.topk(2, dim=1).values
    decided = (top2[:, 0] - top2[:, 1]) > 2 * LOGIT_REL * ref.abs().amax(1)
    same = got.argmax(1) == ref.argmax(1)
    return int(same.sum()), int(decided.sum()), bool(same[decided].all())


def unet_gates(unet_sd, imgs, masks, boxes, device, seg_logits=None, padding=30):
    """Stage-1 gates.  imgs float [B,3,224,224] (NumPy); masks u8 [B,224,224], boxes i32 [B,4] from the engine."""
    masks = np.asarray(masks)
    boxes = np.asarray(boxes)
    ref_logits = oracle_unet_logits(unet_sd, imgs, device)
    ref_masks = unet_ref.mask_from_logits(ref_logits)[:, 0].numpy()
    B = masks.shape[0]
    same_mask = (masks == ref_masks).reshape(B, -1).all(1)
    own_boxes = np.array([roi_ref.bbox_from_mask(m, padding) for m in masks], np.int32)
    ref_boxes = np.array([roi_ref.bbox_from_mask(m, padding) for m in ref_masks], np.int32)
    out = {
        "images": int(B),
        "mask_agreement": float((masks == ref_masks).mean()),
        "mask_agreement_min_image": float((masks == ref_masks).reshape(B, -1).mean(1).min()),
        "masks_identical": int(same_mask.sum()),
        "boxes_bit_exact_given_mask": int((boxes == own_boxes).all(1).sum()),
        "boxes_equal_reference": int((boxes == ref_boxes).all(1).sum()),
        "same_mask_boxes_equal_reference": bool((boxes[same_mask] == ref_boxes[same_mask]).all()),
        "mask_foreground_fraction": float(ref_masks.mean()),
    }
    if seg_logits is not None:
        d = (torch.as_tensor(seg_logits).float().cpu() - ref_logits)
        out["seg_logit_rel_fro"] = float(d.norm() / ref_logits.norm())
        out["seg_logit_max_err_over_scale"] = float(d.abs().max() / ref_logits.abs().max())
        out["seg_logit_mean_abs_err"] = float(d.abs().mean())
    out["ok"] = bool(out["mask_agreement"] >= MASK_AGREEMENT and out["boxes_bit_exact_given_mask"] == B and
                     out["same_mask_boxes_equal_reference"])
    return out, ref_boxes


def pipeline_gates(unet_sd, gnet_sd, imgs, masks, boxes, cls_logits, device, crops_u8=None, seg_logits=None,
                   padding=30):
    """All three gates for a batch.  imgs: float32 [B,3,224,224] NumPy (what the UNet saw); masks / boxes / cls_logits
    (/ crops_u8 [B,224,224,3], seg_logits): engine outputs (tensors or arrays).  Returns a JSON-able dict with "ok"."""
    imgs = np.asarray(imgs, np.float32)
    masks = torch.as_tensor(masks).cpu().numpy()
    boxes = torch.as_tensor(boxes).cpu().numpy()
    cls = torch.as_tensor(cls_logits).float().cpu()
    out, ref_boxes = unet_gates(unet_sd, imgs, masks, boxes, device, seg_logits, padding)
    B = masks.shape[0]
    # reference crop/resize chain of the engine's boxes (== the reference path wherever the box equals the reference's)
    ref_u8 = np.stack([roi_ref.roi_crop_resize_u8(imgs[i], tuple(int(v) for v in boxes[i])) for i in range(B)])
    if crops_u8 is not None:
        got_u8 = torch.as_tensor(crops_u8).cpu().numpy()
        out["crops_bit_exact"] = int((got_u8 == ref_u8).reshape(B, -1).all(1).sum())
    crops = np.transpose(ref_u8, (0, 3, 1, 2)).astype(np.float32) / np.float32(255)
    ref_cls = oracle_googlenet_logits(gnet_sd, crops, device)
    rel = logit_rel_err(cls, ref_cls)
    same_box = (boxes == ref_boxes).all(1)
    out["cls_logit_rel_err_max"] = float(rel.max())
    out["cls_logit_rel_err_max_reference_box"] = float(rel[same_box].max()) if same_box.any() else None
    out["cls_argmax_equal"], out["cls_decided_images"], decided_ok = argmax_gate(cls, ref_cls)
    out["cls_argmax_equal_on_decided"] = decided_ok   # decided: reference top-2 margin > 2 * 1e-2 * logit scale
    out["ok"] = bool(out["ok"] and out["cls_logit_rel_err_max"] <= LOGIT_REL and decided_ok and
                     out.get("crops_bit_exact", B) == B)
    return out


def pil_front_end(src_u8, size=224):
    """Reference front-end of uint8 HWC sources (CDDataAugmentation.transform, data_utils.py:146-147): Pillow bilinear
    resize + to_tensor -> float32 [B,3,S,S].  Pillow itself (third-party, present in the image) is the reference here;
    oracle/roi_ref.pil_resize_bilinear_u8 restates it bit-exactly (tests/test_oracle_roi.py)."""
    from PIL import Image
    out = np.empty((len(src_u8), 3, size, size), np.float32)
    for i, a in enumerate(src_u8):
        r = np.asarray(Image.fromarray(np.ascontiguousarray(a)).resize((size, size), Image.BILINEAR))
        out[i] = np.transpose(r, (2, 0, 1)).astype(np.float32) / np.float32(255)
    return out
