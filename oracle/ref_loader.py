"""Load the UNMODIFIED reference hot-path functions from /root/reference (test infrastructure).

This file is part of the ORACLE side of the repo.  It is only usable in the build
container (where /root/reference is mounted read-only); nothing on the GPU box may
import it.  Its single purpose is to (1) prove `oracle/dcd_oracle.py` equal to the
reference and (2) generate the committed fixtures under tests/golden/ via
`oracle/make_golden.py`.

Loading recipe: the two hot-path modules import a few packages
that are absent here (matplotlib, the DGDE `data` package, GMW `evaluation`) but never
touch them on the hot path, so they are replaced by empty stub modules.  No reference
file is modified or copied.
"""
import importlib
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DCD_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "DGDE", "model", "anno_encoder.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def _stub_matplotlib():
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl = _stub("matplotlib")
        plt = _stub("matplotlib.pyplot")
        mpl.pyplot = plt


_DGDE = None
_GMW = None


def load_dgde_anno_encoder():
    """Return an `Anno_Encoder` instance (DGDE/model/anno_encoder.py) built without cfg.

    `decode_pairs_kpts_depth` (:326-390) and `get_up` (:313-324) use no instance state.
    """
    global _DGDE
    if _DGDE is not None:
        return _DGDE
    _stub_matplotlib()
    _stub("data")
    _stub("data.datasets")
    _stub("data.datasets.kitti_utils", convertAlpha2Rot=lambda *a, **k: None)
    path = os.path.join(REFERENCE_ROOT, "DGDE", "model", "anno_encoder.py")
    spec = importlib.util.spec_from_file_location("_dcd_ref_anno_encoder", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _DGDE = object.__new__(mod.Anno_Encoder)
    return _DGDE


def load_gmw():
    """Return (main_module, GMW_class) of the reference GMW package.

    GMW/main.py:373 compute_z, :364 compute_reg_loss; GMW/model/model.py:103 GMW.
    """
    global _GMW
    if _GMW is not None:
        return _GMW
    _stub_matplotlib()
    _stub("evaluation", evaluate_python=lambda *a, **k: None)
    import numpy.lib.npyio as npyio
    if not hasattr(npyio, "zipfile_factory"):
        npyio.zipfile_factory = None  # name removed in numpy 2; GMW/main.py:18 imports it, never uses it
    gmw_root = os.path.join(REFERENCE_ROOT, "GMW")
    saved_argv = sys.argv
    sys.argv = [saved_argv[0] if saved_argv else "prog"]  # two argparse parsers read sys.argv
    sys.path.insert(0, gmw_root)
    try:
        # the reference packages are called `model`, `lib`, `utilities`: import under their own names
        main = importlib.import_module("main")
        model_mod = importlib.import_module("model.model")
    finally:
        sys.argv = saved_argv
        sys.path.remove(gmw_root)
    _GMW = (main, model_mod.GMW)
    return _GMW


def new_gmw_model(seed: int):
    """Instantiate the reference GMW module with seeded random weights (CPU, FP32)."""
    import torch
    _, GMW = load_gmw()
    saved_argv = sys.argv
    sys.argv = [saved_argv[0] if saved_argv else "prog"]
    try:
        torch.manual_seed(seed)
        model = GMW(None)
    finally:
        sys.argv = saved_argv
    return model.eval()
