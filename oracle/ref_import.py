"""ORACLE helper (build container only): import the REAL reference modules from /root/reference.

/root/reference does not exist on the GPU box, so nothing under tests/ -m gpu, smoke() or bench.py calls
this; it is used by oracle/make_golden.py (golden-vector generation) and by the optional CPU test that
cross-checks the restatement against the live reference when the directory is present.

The reference needs two stubs to import here: `torchsummary` (basicUnet.py:8) and
`matplotlib.pyplot` (roi.py:7).  `分割` and `分类` both define top-level packages `nets`/`util`, so each is
imported under a private module-name prefix by temporarily putting its directory first on sys.path and
evicting the previously imported `nets`/`util` modules.
"""
import importlib
import os
import sys
import types

REF_ROOT = "/root/reference"
SEG_DIR = os.path.join(REF_ROOT, "分割")
CLS_DIR = os.path.join(REF_ROOT, "分类")


def available():
    return os.path.isdir(SEG_DIR) and os.path.isdir(CLS_DIR)


def _install_stubs():
    if "torchsummary" not in sys.modules:
        m = types.ModuleType("torchsummary")
        m.summary = lambda *a, **k: None
        sys.modules["torchsummary"] = m
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mp = types.ModuleType("matplotlib")
        pp = types.ModuleType("matplotlib.pyplot")
        mp.pyplot = pp
        sys.modules["matplotlib"] = mp
        sys.modules["matplotlib.pyplot"] = pp


def _import_from(directory, names):
    _install_stubs()
    saved = {k: v for k, v in sys.modules.items() if k == "nets" or k.startswith("nets.") or k == "util"
             or k.startswith("util.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, directory)
    try:
        mods = [importlib.import_module(n) for n in names]
    finally:
        sys.path.remove(directory)
        for k in [k for k in sys.modules if k == "nets" or k.startswith("nets.") or k == "util"
                  or k.startswith("util.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return mods


def reference_unet_class():
    (m,) = _import_from(SEG_DIR, ["nets.basicUnet"])
    return m.UNetTaskAligWeight


def reference_unet_cls_class():
    """-> the classifier-head UNetTaskAligWeight of 分类/nets/basicUnet.py:369-436 (forward returns cl_out [B,1])."""
    (m,) = _import_from(CLS_DIR, ["nets.basicUnet"])
    return m.UNetTaskAligWeight


def reference_roi():
    """-> (process_and_augment_roi, CDDataAugmentation) from 分类/util."""
    roi, du = _import_from(CLS_DIR, ["util.roi", "util.data_utils"])
    return roi.process_and_augment_roi, du.CDDataAugmentation
