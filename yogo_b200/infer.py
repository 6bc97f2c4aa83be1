"""Inference post-processing entry points of /root/reference/yogo/infer.py:60-124."""
from __future__ import annotations

from typing import Optional

import torch

from .utils.prediction_formatting import format_preds_batch


def get_prediction_class_counts(
    batch_preds: torch.Tensor,
    obj_thresh=0.5,
    iou_thresh=0.5,
    min_class_confidence_threshold: float = 0,
) -> torch.Tensor:
    """Count predictions per argmax class over a batch (infer.py:60-87).  The reference loops
    over images on the CPU; here threshold + NMS + counting is one launch and the (C,) int64
    result is returned on the CPU like the reference's accumulator."""
    _, _, _, counts = format_preds_batch(
        batch_preds, obj_thresh=obj_thresh, iou_thresh=iou_thresh,
        min_class_confidence_threshold=min_class_confidence_threshold,
    )
    return counts.cpu()


def count_cells_for_formatted_preds(
    formatted_class_predictions: torch.Tensor,
    min_confidence_threshold: Optional[float] = None,
) -> torch.Tensor:
    """infer.py:90-124 - host-side helper on an already formatted (N, num_classes) tensor; tiny
    integer bookkeeping, kept in torch."""
    if not len(formatted_class_predictions.shape) == 2:
        raise ValueError(
            "expected formatted_class_predictions to be shape (N, num_classes); "
            f"got {formatted_class_predictions.shape}"
        )
    if min_confidence_threshold is not None:
        if min_confidence_threshold < 0 or min_confidence_threshold > 1:
            raise ValueError(f"min_confidence_threshold should be between 0 and 1; is {min_confidence_threshold}")
    else:
        min_confidence_threshold = 0
    _, n_classes = formatted_class_predictions.shape
    values, indices = formatted_class_predictions.max(dim=1)
    mask = values > min_confidence_threshold
    return torch.nn.functional.one_hot(indices[mask], num_classes=n_classes).sum(dim=0)
