"""Import the reference's own hot-path modules, unmodified, from /root/reference.

Only usable in the authoring container (the GPU box has no /root/reference);
used by ``tests/golden/make_golden.py`` to generate the committed fixtures and
by the ``-m "not gpu"`` tests that compare the oracle with the live reference
when the tree is present.

The reference's ``modules/pose_estimator.py:1-2`` imports ``matplotlib`` and
``onnxruntime`` at module top although the decode/geometry staticmethods only
use numpy; both are absent here, so empty module objects are planted in
``sys.modules`` before the import.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("HBP_REFERENCE_ROOT", "/root/reference")
REF_PKG = os.path.join(REF_ROOT, "human_body_length_est")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_PKG, "modules", "pose_estimator.py"))


def load():
    """Return (pose_estimator_module, onnx_utils_module, utils_module)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    import importlib.machinery
    import torch        # noqa: F401  (imported before the stubs are planted:
    import torchvision  # noqa: F401   torch._dynamo probes sys.modules specs)
    for name in ("matplotlib", "matplotlib.pyplot", "onnxruntime"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                stub = types.ModuleType(name)
                stub.__spec__ = importlib.machinery.ModuleSpec(name, None)
                sys.modules[name] = stub
    if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF_PKG not in sys.path:
        sys.path.insert(0, REF_PKG)
    import modules.pose_estimator as pe      # noqa: E402
    import modules.onnx_utils as ou          # noqa: E402
    import modules.utils as ut               # noqa: E402
    return pe, ou, ut
