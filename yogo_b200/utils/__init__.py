from .prediction_formatting import format_preds, format_preds_batch, split_formatted  # noqa: F401
