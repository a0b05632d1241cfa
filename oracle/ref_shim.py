"""TEST INFRASTRUCTURE ONLY — imports the UNMODIFIED reference modules from /root/reference.

Only usable where /root/reference exists (the build container, not the GPU box).  Used by
`oracle/make_golden.py` to generate `tests/golden/*.pt` and by `tests/test_oracle_vs_reference.py`
to pin the restated oracle (`oracle/sradsgan_oracle.py`) against the reference's own classes
(`SRADSGAN/model/sradsgan.py`: GeneratorResNet :420, Discriminator :470, GANLoss :35).

`model/sradsgan.py` imports, at module scope, packages that are absent here (skimage.measure.compare_*,
matplotlib, tensorflow, sewar, imageio, thop, h5py — SURVEY.md §8c); they are only used for logging /
metrics, never for the arithmetic of the hot path, so inert stubs are registered before the import.
"""
import importlib
import importlib.machinery
import os
import sys
import types

REF_ROOT = os.environ.get("SRADSGAN_REFERENCE", "/root/reference/SRADSGAN")

_STUBS = [
    "skimage", "skimage.measure", "skimage.color", "skimage.transform", "skimage.io", "skimage.metrics",
    "matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "matplotlib.image", "matplotlib.gridspec",
    "tensorflow", "sewar", "sewar.full_ref", "imageio", "thop", "h5py", "cv2",
]


_DATA_FACTORIES = {"data.data": ["get_training_datasets", "get_test_datasets", "get_RGB_trainDataset", "get_RGB_testDataset"]}
_MISSING = {"model.srgan": _DATA_FACTORIES, "model.ndsrgan": _DATA_FACTORIES}


class _Anything(types.ModuleType):
    """Module whose every attribute is a callable returning None (never reached by the hot path)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)

        def _stub(*a, **k):  # inert: module-scope calls such as plt.switch_backend('agg') must pass
            return None

        return _stub


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "model"))


_cached = {}


def load_reference(module="model.sradsgan"):
    """Returns the reference's `model.sradsgan` module (or another `model.*` module, e.g. `model.edsr`), with
    `utils.utils` as attribute .srutils_mod."""
    if module in _cached:
        return _cached[module]
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    import torchvision  # noqa: F401  (must be imported before cv2/skimage stubs are registered)
    for name in _STUBS:
        try:
            importlib.import_module(name)
            continue
        except Exception:
            pass
        m = _Anything(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        m.__path__ = []
        sys.modules[name] = m
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, m)
    saved = {k: sys.modules.get(k) for k in ("model", "utils", "data")}
    for k in list(sys.modules):
        if k in ("model", "utils", "data") or k.startswith(("model.", "utils.", "data.")):
            del sys.modules[k]
    sys.path.insert(0, REF_ROOT)
    try:
        # model/srgan.py:34 imports names its own data/data.py does not define (the sibling cannot be imported in the reference tree
        # as shipped); they are dataset factories, never reached by the arithmetic: inert stand-ins on the reference's module
        for host, names in _MISSING.get(module, {}).items():
            hm = importlib.import_module(host)
            for n in names:
                if not hasattr(hm, n):
                    setattr(hm, n, lambda *a, **k: None)
        mod = importlib.import_module(module)
        mod.srutils_mod = importlib.import_module("utils.utils")
    finally:
        sys.path.remove(REF_ROOT)
    # keep the reference's packages importable under private names only, so that our own
    # top-level names are not shadowed for the rest of the process
    for k in list(sys.modules):
        if k in ("model", "utils", "data") or k.startswith(("model.", "utils.", "data.")):
            sys.modules["_sradsgan_ref_." + k] = sys.modules.pop(k)
    for k, v in saved.items():
        if v is not None:
            sys.modules[k] = v
    _cached[module] = mod
    return mod
