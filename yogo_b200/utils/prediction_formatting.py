"""``format_preds`` - same interface as /root/reference/yogo/utils/prediction_formatting.py:23-93,
plus ``format_preds_batch`` which post-processes a whole batch in one kernel launch
(csrc/nms.cu) instead of one Python iteration + 8 ATen/torchvision calls per image."""
from __future__ import annotations

from typing import List, Literal, Tuple, get_args

import torch

from .. import _lib as L

BoxFormat = Literal["xyxy", "cxcywh"]


def format_preds_batch(
    preds: torch.Tensor,
    obj_thresh: float = 0.5,
    iou_thresh: float = 0.5,
    box_format: BoxFormat = "cxcywh",
    min_class_confidence_threshold: float = 0.0,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """preds (B, 5+C, Sy, Sx) on the GPU -> (rows (B, Sy*Sx, 5+C), keep_count (B,) int32,
    keep_index (B, Sy*Sx) int32, class_counts (C,) int64).  Image b's formatted predictions
    are ``rows[b, :keep_count[b]]`` in the reference's output order; everything stays on the
    device and nothing synchronises."""
    if preds.ndim != 4:
        raise ValueError(f"expected batched predictions (B, pred_shape, Sy, Sx), got {tuple(preds.shape)}")
    if box_format not in get_args(BoxFormat):
        raise ValueError(f"invalid box format {box_format}; valid box formats are {get_args(BoxFormat)}")
    L.require_cuda(preds, "predictions")
    lib = L.lib()
    B, D, Sy, Sx = preds.shape
    p = preds.detach().contiguous().float()
    dev = p.device
    keep_count = torch.empty(B, dtype=torch.int32, device=dev)
    rows = torch.empty((B, Sy * Sx, D), dtype=torch.float32, device=dev)
    keep_index = torch.empty((B, Sy * Sx), dtype=torch.int32, device=dev)
    counts = torch.empty(D - 5, dtype=torch.int64, device=dev)
    nbytes = lib.yg_format_preds_workspace(B, D - 5, Sy, Sx)
    ws = L.workspace.get("nms", nbytes, dev)
    L.check(lib.yg_format_preds_batch(p.data_ptr(), B, D - 5, Sy, Sx, float(obj_thresh), float(iou_thresh),
                                      1 if box_format == "xyxy" else 0, float(min_class_confidence_threshold),
                                      keep_count.data_ptr(), rows.data_ptr(), keep_index.data_ptr(),
                                      counts.data_ptr(), ws.data_ptr(), nbytes, L.stream()))
    return rows, keep_count, keep_index, counts


def format_preds(
    pred: torch.Tensor,
    obj_thresh: float = 0.5,
    iou_thresh: float = 0.5,
    box_format: BoxFormat = "cxcywh",
    min_class_confidence_threshold: float = 0.0,
) -> torch.Tensor:
    """Unbatched (pred_shape, Sy, Sx) -> (n, pred_shape) rows after objectness threshold,
    NMS (rows in descending NMS-score order) and the class-confidence filter."""
    if len(pred.shape) != 3:
        raise ValueError(
            "argument to format_pred should be unbatched result - "
            f"shape should be (pred_shape, Sy, Sx), got {pred.shape}"
        )
    elif box_format not in get_args(BoxFormat):
        raise ValueError(f"invalid box format {box_format}; valid box formats are {get_args(BoxFormat)}")
    rows, keep_count, _, _ = format_preds_batch(
        pred.unsqueeze(0), obj_thresh, iou_thresh, box_format, min_class_confidence_threshold
    )
    n = int(keep_count.item())
    return rows[0, :n].clone()


def split_formatted(rows: torch.Tensor, keep_count: torch.Tensor) -> List[torch.Tensor]:
    """Per-image views of a ``format_preds_batch`` result (one device->host read of the counts)."""
    counts = keep_count.tolist()
    return [rows[b, :n] for b, n in enumerate(counts)]
