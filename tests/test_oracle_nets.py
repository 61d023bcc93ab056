"""CPU: pin the network oracles.

* UNet restatement (oracle/unet_ref.py) vs golden logits of the imported reference UNetTaskAligWeight on the
  seeded procedural weights (tests/golden/unet_golden.npz);
* GoogLeNet restatement vs torchvision's module (the reference's own dependency, present in the image) and vs
  tests/golden/googlenet_golden.npz;
* the shells' state_dict key/shape layout vs the reference's (tests/golden/*_state_keys.json);
* when /root/reference is mounted (build container), the live reference as well."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import fixtures, googlenet_ref, ref_import, unet_ref

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_unet_shell_keys_match_reference():
    with open(os.path.join(GOLD, "unet_state_keys.json")) as f:
        ref = json.load(f)
    tmpl = fixtures.unet_template()
    assert list(tmpl.keys()) == list(ref.keys())
    assert all(list(tmpl[k].shape) == ref[k] for k in ref)
    assert len(ref) == 287


def test_googlenet_shell_keys_match_reference():
    with open(os.path.join(GOLD, "googlenet_state_keys.json")) as f:
        ref = json.load(f)
    import ugnet_b200  # noqa: F401
    from ugnet_b200.googlenet import GoogLeNetClassifier
    sd = GoogLeNetClassifier(6).state_dict()
    assert list(sd.keys()) == list(ref.keys())
    assert all(list(sd[k].shape) == ref[k] for k in ref)
    assert len(ref) == 344 and not any(k.startswith("googlenet.aux") for k in ref)


def test_unet_oracle_matches_reference_golden():
    gold = np.load(os.path.join(GOLD, "unet_golden.npz"))["logits"]
    sd = fixtures.procedural_state(fixtures.unet_template(), seed=7)
    imgs, _, _ = fixtures.synth_images(2, seed=99)
    with torch.no_grad():
        out = unet_ref.unet_forward(sd, torch.from_numpy(imgs)).numpy()
    scale = np.abs(gold).max()
    assert np.abs(out - gold).max() <= 2e-5 * scale, np.abs(out - gold).max() / scale


def test_unet_cls_head_oracle_matches_reference_golden():
    """分类/nets/basicUnet.py:369-436 (classifier-head variant, same state_dict): golden cl_out of the imported class."""
    gold = np.load(os.path.join(GOLD, "unet_cls_golden.npz"))["cl_out"]
    sd = fixtures.procedural_state(fixtures.unet_template(), seed=7)
    imgs, _, _ = fixtures.synth_images(2, seed=99)
    with torch.no_grad():
        out = unet_ref.unet_cls_forward(sd, torch.from_numpy(imgs)).numpy()
    assert out.shape == gold.shape == (2, 1)
    assert np.abs(out - gold).max() <= 2e-5 * np.abs(gold).max()


def test_googlenet_oracle_matches_torchvision_and_golden():
    import torchvision
    gold = np.load(os.path.join(GOLD, "googlenet_golden.npz"))["logits"]
    gsd = fixtures.procedural_state(fixtures.googlenet_template(), seed=11)
    imgs, masks, _ = fixtures.synth_images(4, seed=5)
    crops = torch.from_numpy(fixtures.roi_crops_from_masks(imgs, masks))
    with torch.no_grad():
        out = googlenet_ref.googlenet_forward(gsd, crops).numpy()
    assert np.abs(out - gold).max() <= 1e-4 * max(1.0, np.abs(gold).max())
    net = torchvision.models.googlenet(weights=None, aux_logits=False, transform_input=True, init_weights=False)
    net.fc = torch.nn.Linear(1024, 6)
    net.load_state_dict({k[len("googlenet."):]: v for k, v in gsd.items()})
    net.eval()
    with torch.no_grad():
        tv = net(crops).numpy()
    assert np.abs(out - tv).max() <= 1e-4 * max(1.0, np.abs(tv).max())


@pytest.mark.skipif(not ref_import.available(), reason="reference tree only exists in the build container")
def test_unet_oracle_matches_live_reference():
    Ref = ref_import.reference_unet_class()
    ref = Ref(n_channels=3, n_classes=1).eval()
    sd = fixtures.procedural_state(fixtures.unet_template(), seed=3)
    ref.load_state_dict(sd, strict=True)
    imgs, _, _ = fixtures.synth_images(1, seed=17)
    x = torch.from_numpy(imgs)
    with torch.no_grad():
        a, b = ref(x), unet_ref.unet_forward(sd, x)
    assert (a - b).abs().max() <= 2e-5 * a.abs().max()
