"""CPU: host-side lowering logic with a stub engine (no kernels run).

* every workspace tensor an op descriptor points into stays alive for as long as the compiled program does
  (the descriptors hold raw pointers only; round-1 advisor finding);
* plans are cached with a bound (least recently used dropped);
* weight packing happens on the host and ends in one device buffer;
* the nn.Module shells re-pack after in-place weight edits and never copy / pickle their runner.
"""
import copy
import gc
import pickle
import weakref

import pytest
import torch

import ugnet_b200  # noqa: F401
from ugnet_b200 import engine as E
from ugnet_b200 import lower


class _FakeProgram:
    def __init__(self, descs, keepalive):
        self.descs, self.keepalive = list(descs), list(keepalive)


class _FakeEngine:
    launch_count = 0

    def program(self, descs, keepalive=()):
        return _FakeProgram(descs, keepalive)


@pytest.fixture()
def fake_engine(monkeypatch):
    monkeypatch.setattr(E.Engine, "get", classmethod(lambda cls, device=0: _FakeEngine()))


def _tracking(builder, refs):
    orig = builder.buf

    def buf(*shape, dtype=torch.bfloat16):
        t = orig(*shape, dtype=dtype)
        refs.append(weakref.ref(t))
        return t
    builder.buf = buf


def _unet_sd():
    from ugnet_b200.nets import UNetTaskAligWeight
    torch.manual_seed(0)
    return UNetTaskAligWeight(3, 1).state_dict()


def _gnet_sd():
    from ugnet_b200.googlenet import GoogLeNetClassifier
    torch.manual_seed(0)
    return GoogLeNetClassifier(6).state_dict()


@pytest.mark.parametrize("head", ["seg", "cls"])
def test_unet_workspace_outlives_emission(fake_engine, head):
    r = lower.UNetRunner(_unet_sd(), "cpu", head=head)
    refs = []
    _tracking(r, refs)
    ws = r.plan(1)
    gc.collect()
    assert len(refs) > (40 if head == "seg" else 20)
    dead = [i for i, w in enumerate(refs) if w() is None]
    assert not dead, f"{len(dead)} of {len(refs)} workspace tensors were freed while the program still points at them"
    held = {id(t) for t in ws["program"].keepalive}
    assert all(id(w()) in held for w in refs)


def test_googlenet_and_pipeline_workspace_outlive_emission(fake_engine):
    g = lower.GoogLeNetRunner(_gnet_sd(), "cpu")
    refs = []
    _tracking(g, refs)
    g.plan(1, "u8")
    gc.collect()
    assert len(refs) > 30 and all(w() is not None for w in refs)

    pipe = lower.PipelineRunner(_unet_sd(), _gnet_sd(), "cpu", micro_batch=1)
    refs = []
    _tracking(pipe.unet, refs)
    _tracking(pipe.gnet, refs)
    ws = pipe.plan(2)                      # two micro-batches share one UNet workspace
    gc.collect()
    assert all(w() is not None for w in refs)
    held = {id(t) for t in ws["program"].keepalive}
    assert all(id(w()) in held for w in refs)


def test_plan_cache_is_bounded(fake_engine, monkeypatch):
    r = lower.GoogLeNetRunner(_gnet_sd(), "cpu")
    r.plans.limit = 2
    a = r.plan(1, "u8")
    r.plan(2, "u8")
    assert r.plan(1, "u8") is a            # hit: becomes most recently used
    r.plan(3, "u8")                        # evicts batch 2, the least recently used
    assert set(r.plans) == {(1, "u8"), (3, "u8")}
    ref = weakref.ref(a["program"])
    del a
    r.plan(4, "u8")
    r.plan(5, "u8")
    gc.collect()
    assert ref() is None, "an evicted plan must release its program and workspace"

    pipe = lower.PipelineRunner(_unet_sd(), _gnet_sd(), "cpu", micro_batch=2)
    pipe.plans.limit = 1
    pipe.plan(1)
    pipe.plan(2)
    assert set(pipe._pools) == {2}, "the shared workspace of an evicted micro-batch size is dropped with its plan"


def test_packing_is_host_side_and_lands_in_one_buffer(fake_engine):
    r = lower.UNetRunner(_unet_sd(), "cpu")
    base, size = r.w_blob.data_ptr(), r.w_blob.numel()
    seen = []
    lower._map_tensors(r.w, lambda t: seen.append(t) or t)
    assert len(seen) > 100
    for t in seen:
        assert base <= t.data_ptr() < base + size and t.data_ptr() % 16 == 0
    w = r.w["down1.0"]
    assert w["w"].dtype == torch.bfloat16 and w["w"].shape == (128, 9 * 64) and w["scale"].dtype == torch.float32


def test_shell_repacks_after_in_place_weight_edits():
    from ugnet_b200.nets import UNetTaskAligWeight
    m = UNetTaskAligWeight(3, 1).eval()
    k0 = m._fingerprint()
    assert m._fingerprint() == k0
    with torch.no_grad():
        m.outc.weight.mul_(2.0)
    k1 = m._fingerprint()
    assert k1 != k0, "an in-place parameter edit must change the fingerprint"
    with torch.no_grad():
        m.inc.conv.weight.copy_(torch.zeros_like(m.inc.conv.weight))
    k2 = m._fingerprint()
    assert k2 != k1
    m.fc1 = torch.nn.Linear(512, 256)              # a replaced submodule has new storage
    assert m._fingerprint() != k2
    # writes through `.data` bypass autograd's version counter: those (only) need an explicit invalidate()
    m._runner = object()
    m.invalidate()
    assert m._runner is None
    m._runner, m._runner_key = object(), k0        # a stale runner from before the edits
    with pytest.raises(RuntimeError):              # CPU module: runner() drops the stale one, then refuses to build
        m.runner()
    assert m._runner is None


def test_shell_copies_never_carry_the_runner():
    from ugnet_b200.googlenet import GoogLeNetClassifier
    from ugnet_b200.nets import UNetTaskAligWeight
    for m in (UNetTaskAligWeight(3, 1).eval(), GoogLeNetClassifier(6).eval()):
        class _Unpicklable:
            def __reduce__(self):
                raise TypeError("engine handles cannot be pickled")
        m._runner, m._runner_key = _Unpicklable(), m._fingerprint()
        c = copy.deepcopy(m)
        assert c._runner is None and m._runner is not None
        p = pickle.loads(pickle.dumps(m))
        assert p._runner is None
        assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), p.state_dict().values()))
