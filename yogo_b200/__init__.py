"""yogo_b200 - B200-native implementation of the YOGO hot path behind the reference's Python API."""
from .model import YOGO  # noqa: F401
from .model_defns import MODELS, get_model_func, register_model  # noqa: F401
from .yogo_loss import YOGOLoss  # noqa: F401
from .utils import format_preds, format_preds_batch  # noqa: F401
from .infer import get_prediction_class_counts, count_cells_for_formatted_preds  # noqa: F401

__all__ = [
    "YOGO", "YOGOLoss", "MODELS", "get_model_func", "register_model", "format_preds", "format_preds_batch",
    "get_prediction_class_counts", "count_cells_for_formatted_preds",
]
