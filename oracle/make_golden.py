"""Generate tests/golden/*.npz by running the IMPORTED reference (build container only; the script is
committed with the vectors it made).  Run from the repo root:  python -m oracle.make_golden

  unet_golden.npz       reference UNetTaskAligWeight(3,1) logits for 2 synthetic images with the seeded
                        procedural weights (oracle.fixtures.procedural_state, seed 7) + the reference's
                        state_dict key list and shapes
  unet_cls_golden.npz   the classifier-head UNetTaskAligWeight of 分类/nets/basicUnet.py on the same weights and images
  googlenet_golden.npz  torchvision GoogLeNet (built as 分类/test.py:64-73 builds it) logits for 4 crops with
                        procedural weights (seed 11)
  roi_golden.npz        reference process_and_augment_roi outputs (roi tensor as uint8, box) for hand-made
                        masks, obtained by giving the reference function a stub model that returns prescribed
                        logits
"""
import json
import os

import numpy as np
import torch

from . import fixtures, ref_import

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class _StubSeg(torch.nn.Module):
    """Returns logits whose sigmoid > 0.5 exactly on a prescribed mask."""

    def __init__(self, mask):
        super().__init__()
        self.mask = mask

    def forward(self, x):
        return (torch.from_numpy(self.mask).float() * 8.0 - 4.0)[None, None]


def main():
    assert ref_import.available(), "reference tree not mounted"
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)

    # ---- UNet
    RefUNet = ref_import.reference_unet_class()
    ref = RefUNet(n_channels=3, n_classes=1).eval()
    ref_sd = ref.state_dict()
    tmpl = fixtures.unet_template()
    assert list(tmpl.keys()) == list(ref_sd.keys()), "shell key order differs from the reference"
    assert all(tuple(tmpl[k].shape) == tuple(ref_sd[k].shape) for k in tmpl)
    sd = fixtures.procedural_state(tmpl, seed=7)
    ref.load_state_dict(sd, strict=True)
    imgs, _, _ = fixtures.synth_images(2, seed=99)
    with torch.no_grad():
        logits = ref(torch.from_numpy(imgs)).numpy()
    np.savez_compressed(os.path.join(OUT, "unet_golden.npz"), logits=logits.astype(np.float32))
    with open(os.path.join(OUT, "unet_state_keys.json"), "w") as f:
        json.dump({k: list(v.shape) for k, v in ref_sd.items()}, f, indent=0)
    print("unet golden", logits.shape, float(np.abs(logits).mean()))

    # ---- classifier-head variant (分类/nets/basicUnet.py:369-436): same state_dict, forward -> cl_out [B,1]
    RefCls = ref_import.reference_unet_cls_class()
    refc = RefCls(n_channels=3, n_classes=1).eval()
    assert list(refc.state_dict().keys()) == list(ref_sd.keys()), "classifier-head variant has a different state_dict"
    refc.load_state_dict(sd, strict=True)
    with torch.no_grad():
        cl = refc(torch.from_numpy(imgs)).numpy()
    np.savez_compressed(os.path.join(OUT, "unet_cls_golden.npz"), cl_out=cl.astype(np.float32))
    print("unet cls-head golden", cl.ravel())

    # ---- GoogLeNet (torchvision, as the reference constructs it minus the download)
    import torchvision
    net = torchvision.models.googlenet(weights=None, aux_logits=False, transform_input=True, init_weights=False)
    net.fc = torch.nn.Linear(1024, 6)
    gt = fixtures.googlenet_template()
    gsd = fixtures.procedural_state(gt, seed=11)
    net.load_state_dict({k[len("googlenet."):]: v for k, v in gsd.items()}, strict=True)
    net.eval()
    imgs4, masks4, _ = fixtures.synth_images(4, seed=5)
    crops = fixtures.roi_crops_from_masks(imgs4, masks4)
    with torch.no_grad():
        gl = net(torch.from_numpy(crops)).numpy()
    np.savez_compressed(os.path.join(OUT, "googlenet_golden.npz"), logits=gl.astype(np.float32))
    with open(os.path.join(OUT, "googlenet_state_keys.json"), "w") as f:
        json.dump({k: list(v.shape) for k, v in gsd.items()}, f, indent=0)
    print("googlenet golden", gl)

    # ---- ROI path through the reference's own function
    proc, Aug = ref_import.reference_roi()
    # constructed exactly as the inference script does (分类/test.py:113-116): no colour jitter, no flips
    aug = Aug(img_size=224, ori_size=224, crop=None, p_hflip=0.0, p_vflip=0.0, color_jitter_params=None,
              long_mask=True)
    H = W = 224
    masks = []
    m = np.zeros((H, W), np.uint8); masks.append(m.copy())
    m = np.zeros((H, W), np.uint8); m[0, 0] = 1; masks.append(m)
    m = np.zeros((H, W), np.uint8); m[100:130, 0:5] = 1; masks.append(m)
    m = np.zeros((H, W), np.uint8); m[60:160, 219:224] = 1; masks.append(m)
    m = np.zeros((H, W), np.uint8); m[90:101, 70:150] = 1; masks.append(m)
    masks.append(masks4[0]); masks.append(masks4[1])
    imgs7, _, _ = fixtures.synth_images(len(masks), seed=21)
    # make the channels differ so that the BGR->RGB flip is observable
    imgs7[:, 1] = np.clip(imgs7[:, 1] * 0.8 + 0.1, 0, 1)
    imgs7[:, 2] = np.clip(1.0 - imgs7[:, 2], 0, 1)
    rois = []
    for i, mk in enumerate(masks):
        roi, se = proc(_StubSeg(mk), torch.from_numpy(imgs7[i]), torch.device("cpu"), aug, f"{i}.png")
        rois.append(np.round(roi.numpy() * 255.0).astype(np.uint8))
    np.savez_compressed(os.path.join(OUT, "roi_golden.npz"), masks=np.stack(masks), rois=np.stack(rois),
                        images_seed=np.int64(21))
    print("roi golden", np.stack(rois).shape)


if __name__ == "__main__":
    main()
