"""GPU parity of the whole path against the CPU oracle (fp32) on briefly-trained weights.

Gates (BASELINE.json north_star):
  * segmentation masks >= 99.9 % pixel agreement;
  * bbox / crop indices bit-exact given the same mask;
  * class logits within 1e-2 of the per-image logit scale (max|d| <= 1e-2 * max|ref|), identical argmax.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N_IMG = 8


@pytest.fixture(scope="module")
def unet_sd():
    from oracle import fixtures
    return fixtures.trained_unet_state(device="cuda")


@pytest.fixture(scope="module")
def gnet_sd():
    from oracle import fixtures
    return fixtures.trained_googlenet_state(device="cuda")


@pytest.fixture(scope="module")
def images():
    from oracle import fixtures
    imgs, masks, labels = fixtures.synth_images(N_IMG, seed=2024)
    return imgs, masks, labels


@pytest.fixture(scope="module")
def oracle_unet(unet_sd, images):
    from oracle import unet_ref
    with torch.no_grad():
        logits, inter = unet_ref.unet_forward(unet_sd, torch.from_numpy(images[0]), return_intermediates=True)
    return logits, inter


def test_fixture_is_trained_like(oracle_unet, images):
    from oracle import unet_ref
    mask = unet_ref.mask_from_logits(oracle_unet[0])[:, 0].numpy()
    acc = (mask == images[1]).mean()
    assert acc > 0.98, f"fixture UNet only reaches {acc:.4f} pixel accuracy: not trained-like"
    assert (oracle_unet[0].abs() < 0.1).float().mean() < 0.01


def test_unet_intermediates(engine, unet_sd, images, oracle_unet):
    """Layer-by-layer drift check (bf16 storage): relative Frobenius error of each stage output."""
    from ugnet_b200.lower import UNetRunner
    r = UNetRunner(unet_sd, "cuda:0")
    x = torch.from_numpy(images[0]).cuda()
    r.forward(x)
    ws = r.plan(N_IMG)
    _, inter = oracle_unet
    names = {"x1": "x1", "down1": "x2", "down2": "x3", "down3": "x4", "down4": "out0", "bottleneck": "t",
             "up4": "out1", "up3": "out2", "up2": "out3"}
    for k, rk in names.items():
        got = ws[k].float().cpu().permute(0, 3, 1, 2)
        ref = inter[rk]
        rel = ((got - ref).norm() / ref.norm()).item()
        assert rel < 0.03, f"{k}: relative error {rel:.4f}"


def test_unet_parity(engine, unet_sd, images, oracle_unet):
    from oracle import roi_ref, unet_ref
    from ugnet_b200.lower import UNetRunner
    r = UNetRunner(unet_sd, "cuda:0")
    x = torch.from_numpy(images[0]).cuda()
    logits, mask, boxes = r.forward(x, with_mask_boxes=True)
    ref_logits = oracle_unet[0]
    ref_mask = unet_ref.mask_from_logits(ref_logits)[:, 0].numpy()
    agree = (mask.cpu().numpy() == ref_mask).mean()
    assert agree >= 0.999, f"mask agreement {agree:.5f} < 99.9 %"
    d = (logits.cpu() - ref_logits).abs()
    rel_fro = ((logits.cpu() - ref_logits).norm() / ref_logits.norm()).item()
    # the contract gate for the segmentation stage is the mask (above); the logit map itself is held to the measured
    # bf16 drift with ~2.5x headroom: relative Frobenius error 2.4e-3, worst pixel 0.7 % of the logit range
    assert rel_fro <= 6e-3 and d.max() <= 0.02 * ref_logits.abs().max(), (rel_fro, d.mean().item(), d.max().item())
    ref_boxes = np.array([roi_ref.bbox_from_mask(m) for m in ref_mask], np.int32)
    got_boxes = boxes.cpu().numpy()
    same_mask = [(mask[i].cpu().numpy() == ref_mask[i]).all() for i in range(N_IMG)]
    for i in range(N_IMG):
        # bit-exact given the same mask: the engine's bbox of ITS mask equals the oracle bbox of that mask
        assert tuple(got_boxes[i]) == roi_ref.bbox_from_mask(mask[i].cpu().numpy())
        if same_mask[i]:
            assert tuple(got_boxes[i]) == tuple(ref_boxes[i])
    print(f"mask agreement {agree:.6f}; mean|dlogit| {d.mean():.4f}; max|dlogit|/scale "
          f"{(d.max() / ref_logits.abs().max()).item():.4f}; rel fro {rel_fro:.4f}; boxes equal "
          f"{(got_boxes == ref_boxes).all(1).sum()}/{N_IMG}")


def test_unet_shell_is_drop_in(engine, unet_sd, images, oracle_unet):
    """Reference-style usage: construct, load_state_dict(strict), eval, call (predict.py:112-115,23)."""
    from ugnet_b200.nets.basicUnet_new import UNetTaskAligWeight
    model = UNetTaskAligWeight(n_channels=3, n_classes=1).to("cuda")
    model.load_state_dict(unet_sd)
    model.eval()
    x = torch.from_numpy(images[0][:3]).cuda()
    with torch.no_grad():
        out = model(x)
    assert out.shape == (3, 1, 224, 224) and out.dtype == torch.float32
    ref = oracle_unet[0][:3]
    assert ((torch.sigmoid(out.cpu()) > 0.5) == (torch.sigmoid(ref) > 0.5)).float().mean() >= 0.999
    with pytest.raises(ValueError):
        model(torch.zeros(1, 3, 512, 512, device="cuda"))


def test_unet_cls_head_variant(engine, unet_sd, images):
    """分类/nets/basicUnet.py:369-436 (classifier-head UNetTaskAligWeight): same state_dict, forward ->
    cl_out [B,1].  Checked against the golden cl_out of the IMPORTED reference class (procedural weights, seed 7) and
    against the fp32 oracle on the trained fixture.  Tolerance: the contract's 1e-2 of the output scale (measured 4e-3:
    bf16 activations through 10 conv layers + the transformer block, a mean over 196 tokens, a 512-term dot product)."""
    import os
    from oracle import fixtures, unet_ref
    from ugnet_b200.nets.basicUnet_cls import UNetTaskAligWeight
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "unet_cls_golden.npz"))["cl_out"]
    sd = fixtures.procedural_state(fixtures.unet_template(), seed=7)
    imgs, _, _ = fixtures.synth_images(2, seed=99)
    model = UNetTaskAligWeight(n_channels=3, n_classes=1)
    model.load_state_dict(sd, strict=True)
    model = model.to("cuda").eval()
    with torch.no_grad():
        out = model(torch.from_numpy(imgs).cuda())
    assert out.shape == (2, 1) and out.dtype == torch.float32
    err = np.abs(out.cpu().numpy() - gold).max()
    print(f"cls-head golden: engine {out.cpu().numpy().ravel()} reference {gold.ravel()} err {err:.4f}")
    assert err <= 1e-2 * max(1.0, np.abs(gold).max())
    # trained encoder weights, 8 images, live oracle
    model.load_state_dict(unet_sd)
    x = torch.from_numpy(images[0])
    with torch.no_grad():
        ref = unet_ref.unet_cls_forward({k: v.cpu() for k, v in unet_sd.items()}, x).numpy()
        got = model(x.cuda()).cpu().numpy()
    err = np.abs(got - ref).max()
    print(f"cls-head trained fixture: err {err:.4f} of scale {np.abs(ref).max():.3f}")
    assert err <= 1e-2 * max(1.0, np.abs(ref).max())
    with pytest.raises(RuntimeError):
        model.forward_mask_boxes(x.cuda())


def test_roi_drop_in_matches_oracle(engine, unet_sd, images):
    """process_and_augment_roi (roi.py:12-51): crop pixels bit-exact given the engine's own mask."""
    from oracle import roi_ref
    from ugnet_b200.nets import UNetTaskAligWeight
    from ugnet_b200.util.roi import process_and_augment_roi
    model = UNetTaskAligWeight(3, 1).to("cuda")
    model.load_state_dict(unet_sd)
    model.eval()
    img = torch.from_numpy(images[0][1])
    roi, se_out = process_and_augment_roi(model, img, torch.device("cuda"), None, "1.png")
    assert roi.shape == (3, 224, 224) and se_out.shape == (1, 1, 224, 224)
    mask = (torch.sigmoid(se_out) > 0.5)[0, 0].cpu().numpy().astype(np.uint8)
    ref, _ = roi_ref.roi_tensor(images[0][1], mask)
    assert np.array_equal(np.round(roi.cpu().numpy() * 255).astype(np.uint8), np.round(ref * 255).astype(np.uint8))
    assert np.abs(roi.cpu().numpy() - ref).max() < 1e-6


def _logit_gate(got, ref):
    scale = ref.abs().amax(1, keepdim=True)
    rel = ((got - ref).abs() / scale).amax(1)
    return rel


def test_googlenet_parity(engine, gnet_sd, images):
    from oracle import fixtures, googlenet_ref
    from ugnet_b200.googlenet import GoogLeNetClassifier
    crops = fixtures.roi_crops_from_masks(images[0], images[1])          # float [N,3,224,224], k/255 values
    with torch.no_grad():
        ref = googlenet_ref.googlenet_forward(gnet_sd, torch.from_numpy(crops))
    model = GoogLeNetClassifier(num_classes=6).to("cuda")
    model.load_state_dict(gnet_sd)
    model.eval()
    with torch.no_grad():
        got = model(torch.from_numpy(crops).cuda()).cpu()
    rel = _logit_gate(got, ref)
    assert (rel <= 1e-2).all(), f"logit error / scale per image: {rel.tolist()}"
    assert torch.equal(got.argmax(1), ref.argmax(1))
    # uint8 entry (what the pipeline feeds) gives the same logits
    u8 = torch.from_numpy(np.round(crops * 255).astype(np.uint8)).permute(0, 2, 3, 1).contiguous().cuda()
    got_u8 = model.runner().forward_u8(u8).cpu()
    assert (got_u8 - got).abs().max() < 1e-3
    print(f"googlenet max rel err {rel.max():.5f}; acc vs labels {(got.argmax(1).numpy() == images[2]).mean():.2f}")


def test_pipeline_parity(engine, unet_sd, gnet_sd, images, oracle_unet):
    from oracle import googlenet_ref, roi_ref, unet_ref
    from ugnet_b200.pipeline import TwoStagePipeline
    pipe = TwoStagePipeline({k: v.cuda() for k, v in unet_sd.items()}, {k: v.cuda() for k, v in gnet_sd.items()},
                            micro_batch=4)   # 2 micro-batches
    x = torch.from_numpy(images[0]).cuda()
    masks, boxes, cls = pipe(x)
    ref_mask = unet_ref.mask_from_logits(oracle_unet[0])[:, 0].numpy()
    assert (masks.cpu().numpy() == ref_mask).mean() >= 0.999
    ref_boxes = np.array([roi_ref.bbox_from_mask(m) for m in ref_mask], np.int32)
    crops = np.stack([roi_ref.roi_tensor(images[0][i], ref_mask[i])[0] for i in range(N_IMG)])
    with torch.no_grad():
        ref_cls = googlenet_ref.googlenet_forward(gnet_sd, torch.from_numpy(crops))
    got_masks = masks.cpu().numpy()
    same_mask = np.array([(got_masks[i] == ref_mask[i]).all() for i in range(N_IMG)])
    same_box = (boxes.cpu().numpy() == ref_boxes).all(1)
    ndiff = [(got_masks[i] != ref_mask[i]).sum() for i in range(N_IMG)]
    print(f"pipeline: differing mask pixels per image {ndiff}; boxes equal {same_box.sum()}/{N_IMG}")
    # bbox is a min/max over the mask, so one flipped pixel outside the blob moves it: the gate is "bit-exact
    # given the same mask" — every image whose mask equals the oracle's must have the oracle's box and crop
    assert max(ndiff) <= 50, "a few boundary pixels may flip under bf16, not more"
    assert same_box[same_mask].all()
    for i in range(N_IMG):
        assert tuple(boxes[i].tolist()) == roi_ref.bbox_from_mask(got_masks[i])
    sel = torch.from_numpy(same_box)
    rel = _logit_gate(cls.cpu(), ref_cls)
    assert (rel[sel] <= 1e-2).all(), rel.tolist()
    assert torch.equal(cls.cpu().argmax(1)[sel], ref_cls.argmax(1)[sel])
    # images whose box differs: the classifier must still match the oracle run on the ENGINE's own crop
    for i in np.nonzero(~same_box)[0]:
        crop_i, _ = roi_ref.roi_tensor(images[0][i], got_masks[i])
        with torch.no_grad():
            ref_i = googlenet_ref.googlenet_forward(gnet_sd, torch.from_numpy(crop_i)[None])
        assert _logit_gate(cls.cpu()[i:i + 1], ref_i).item() <= 1e-2
    # determinism: a second run is bit-identical
    m2, b2, c2 = pipe(x)
    assert torch.equal(m2, masks) and torch.equal(b2, boxes) and torch.equal(c2, cls)
    print(f"pipeline: cls rel err (same-box images) {rel[sel].max():.5f}")


def test_full_size_batch_invariance(engine, unet_sd, gnet_sd):
    """BASELINE.json's per-GPU size (256 images, UNet micro-batch 128): size-independent properties instead of an
    oracle run.  (1) every image's result is bit-identical to the one computed in a batch of 8 (images are independent
    end to end; exercises the large-batch tile decode); (2) every box is exactly bbox(mask) of the oracle's integer
    rule; (3) the run is deterministic."""
    from oracle import fixtures, roi_ref
    from ugnet_b200.pipeline import TwoStagePipeline
    usd = {k: v.cuda() for k, v in unet_sd.items()}
    gsd = {k: v.cuda() for k, v in gnet_sd.items()}
    imgs, _, _ = fixtures.synth_images(256, seed=77)
    x = torch.from_numpy(imgs).cuda()
    big = TwoStagePipeline(usd, gsd, micro_batch=128)
    masks, boxes, cls = big(x)
    assert masks.shape == (256, 224, 224) and boxes.shape == (256, 4) and cls.shape == (256, 6)
    m2, b2, c2 = big(x)
    assert torch.equal(m2, masks) and torch.equal(b2, boxes) and torch.equal(c2, cls)
    mh = masks.cpu().numpy()
    for i in range(256):
        assert tuple(boxes[i].tolist()) == roi_ref.bbox_from_mask(mh[i]), i
    assert 0.01 < mh.mean() < 0.5, "fixture masks should be non-trivial"
    small = TwoStagePipeline(usd, gsd, micro_batch=8)
    for s0 in (0, 124, 248):      # first chunk, one straddling the micro-batch boundary, last chunk
        ms, bs, cs = small(x[s0:s0 + 8])
        assert torch.equal(ms, masks[s0:s0 + 8]), f"masks differ from the batch-8 run at {s0}"
        assert torch.equal(bs, boxes[s0:s0 + 8])
        assert torch.equal(cs, cls[s0:s0 + 8]), f"logits differ from the batch-8 run at {s0}"


def test_run_host_pipelined_matches_serial(engine, unet_sd, gnet_sd, images):
    """ug_program_run_host_pipelined (double-buffered H2D on the copy stream) == ug_program_run_host, step by step."""
    from ugnet_b200.lower import PipelineRunner
    pipe = PipelineRunner(unet_sd, gnet_sd, "cuda:0", micro_batch=4)
    B = 4
    ws = pipe.plan(B)
    prog = ws["program"]
    steps = [torch.from_numpy(images[0][i:i + B].copy()).pin_memory() for i in (0, 2, 4, 1, 3)]
    serial = []
    for h_in in steps:
        hm = torch.empty((B, 224, 224), dtype=torch.uint8).pin_memory()
        hc = torch.empty((B, 6), dtype=torch.float32).pin_memory()
        prog.run_host([(ws["x_in"], h_in)], [(hm, ws["mask"]), (hc, ws["cls_logits"])])
        serial.append((hm.clone(), hc.clone()))
    outs = [(torch.empty((B, 224, 224), dtype=torch.uint8).pin_memory(), torch.empty((B, 6)).pin_memory()) for _ in steps]
    for h_in, (hm, hc) in zip(steps, outs):
        prog.run_host_pipelined([(ws["x_in"], h_in)], [(hm, ws["mask"]), (hc, ws["cls_logits"])])
    torch.cuda.synchronize()
    for (m0, c0), (m1, c1) in zip(serial, outs):
        assert torch.equal(m0, m1) and torch.equal(c0, c1)
