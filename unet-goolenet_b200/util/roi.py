"""Drop-in for 分类/util/roi.py: mask -> bbox -> crop -> uint8 -> channel flip -> PIL-exact resize, on the GPU.

`process_and_augment_roi` keeps the reference signature and return value (roi tensor [3,224,224] float32 in
[0,1] as produced by to_tensor, plus the raw UNet logits [1,1,224,224]); the batched form used by the pipeline
is `roi_batch`.  `transform_fn` is accepted for signature compatibility: on the reference's inference path it
is CDDataAugmentation(img_size=224) with every augmentation probability at 0, i.e. exactly the resize +
to_tensor implemented by the crop/resize kernel (data_utils.py:102,146-147); other sizes are honoured through
`transform_fn.img_size`."""
import torch

from .. import engine as E


@torch.no_grad()
def roi_batch(model, images, padding=30, out_size=224):
    """images: float [B,3,224,224] CUDA. Returns (roi_u8 [B,S,S,3] uint8 HWC, boxes i32 [B,4], logits, masks)."""
    logits, masks, boxes = model.forward_mask_boxes(images, padding=padding)
    eng = model.runner().engine
    B, _, H, W = images.shape
    img = images.float().contiguous()
    out = torch.empty((B, out_size, out_size, 3), dtype=torch.uint8, device=images.device)
    eng.run_op(E.CropResizeDesc(img.data_ptr(), boxes.data_ptr(), out.data_ptr(), B, H, W, out_size))
    return out, boxes, logits, masks


@torch.no_grad()
def process_and_augment_roi(model, image, device, transform_fn=None, name=None, padding=30):
    """roi.py:12-51 for one image [3,H,W]: returns (roi_tensor_aug [3,S,S] float32, se_out [1,1,H,W])."""
    size = getattr(transform_fn, "img_size", 224) if transform_fn is not None else 224
    if transform_fn is not None:
        probs = [getattr(transform_fn, a, 0.0) for a in ("p_hflip", "p_vflip", "p_rota", "p_gaussn", "p_gama",
                                                         "p_contr", "p_distortion")]
        if any(p > 0 for p in probs) or getattr(transform_fn, "color_jitter_params", None):
            raise NotImplementedError("only the inference configuration of CDDataAugmentation (test.py:113-116: "
                                      "all p_* = 0, color_jitter_params=None) is on the hot path")
    x = image.unsqueeze(0).to(device)
    model.eval()
    roi_u8, _, logits, _ = roi_batch(model, x, padding=padding, out_size=size)
    roi = roi_u8[0].permute(2, 0, 1).float() / 255.0          # F.to_tensor
    return roi.cpu() if image.device.type == "cpu" else roi, logits
