"""Import the UNMODIFIED reference (czbiohub-sf/yogo) for the comparison arms of bench.py: `baseline/_ref` (installed once with
`pip install --no-index --no-deps --ignore-requires-python --target baseline/_ref <copy of /root/reference>`, git-ignored, travels to
the GPU box) or, in the build container, /root/reference itself.  Modules off the hot path that are absent from this image
(matplotlib, zarr, ruamel.yaml; np.unicode_ of numpy < 2) are stubbed as in SURVEY.md 8c.  Never imported by the product."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_path():
    for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(p, "yogo")):
            return p
    return None


def import_reference():
    """-> dict of the reference's public entry points for the path, or None when no copy of the reference is available."""
    path = reference_path()
    if path is None:
        return None
    import numpy as np

    for name in ("matplotlib", "matplotlib.pyplot", "zarr", "ruamel", "ruamel.yaml"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["ruamel.yaml"], "YAML"):
        sys.modules["ruamel.yaml"].YAML = object
        sys.modules["ruamel"].yaml = sys.modules["ruamel.yaml"]
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(np, "unicode_"):
        np.unicode_ = np.str_
    if path not in sys.path:
        sys.path.insert(0, path)
    import yogo  # noqa: F401
    from yogo.infer import get_prediction_class_counts
    from yogo.model import YOGO
    from yogo.model_defns import get_model_func
    from yogo.utils import format_preds
    from yogo.yogo_loss import YOGOLoss

    return dict(YOGO=YOGO, get_model_func=get_model_func, YOGOLoss=YOGOLoss, format_preds=format_preds,
                get_prediction_class_counts=get_prediction_class_counts, path=path)
