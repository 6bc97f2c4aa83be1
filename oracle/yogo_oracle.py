"""CPU restatement of the YOGO hot path (TEST INFRASTRUCTURE - see oracle/__init__.py).

Everything here runs on host cores only.  The dense arithmetic (3x3/1x1 convolutions,
BatchNorm) is restated with ``torch.nn.functional`` on CPU tensors because that *is* the
reference's algorithm (the reference dispatches ``nn.Conv2d``/``nn.BatchNorm2d``,
/root/reference/yogo/model_defns.py:30-77); the head transform, the loss, its analytic
gradient, threshold+NMS and the class counts are restated per cell in numpy from the
reference sources cited on each function.

Parity pinned by tests/golden/ (generated from the real reference by
tests/golden/make_golden.py) - see tests/test_oracle_golden.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------
# Backbone tables.  One row per block: (cout, ksize, stride, conv_bias, batchnorm, act, p_drop)
# act: "act" = the family's activation (LeakyReLU(0.01) or SiLU), None = no activation.
# Restates /root/reference/yogo/model_defns.py (line ranges per entry).
# ----------------------------------------------------------------------------------------
Row = Tuple[int, int, int, bool, bool, Optional[str], float]


def _eight(c1: int, c2: int, c3: int, c4: int) -> List[Row]:
    # topology shared by base/silu/double/triple/half/quarter (model_defns.py:30-77 etc.)
    return [
        (c1, 3, 2, False, True, "act", 0.0),
        (c2, 3, 1, True, False, "act", 0.05),
        (c3, 3, 2, True, False, "act", 0.10),
        (c4, 3, 1, True, False, "act", 0.15),
        (c4, 3, 2, False, True, "act", 0.0),
        (c4, 3, 1, True, True, "act", 0.0),
        (c4, 3, 1, True, False, "act", 0.0),
    ]


BACKBONES: Dict[str, Tuple[str, List[Row]]] = {
    "base_model": ("lrelu", _eight(16, 32, 64, 128)),  # :30-77
    "silu_model": ("silu", _eight(16, 32, 64, 128)),  # :80-127
    "double_filters": ("lrelu", _eight(32, 64, 128, 256)),  # :130-177
    "triple_filters": ("lrelu", _eight(48, 96, 192, 384)),  # :180-227
    "half_filters": ("lrelu", _eight(8, 16, 32, 64)),  # :230-277
    "quarter_filters": ("lrelu", _eight(4, 8, 16, 32)),  # :280-327
    "depth_ver_0": (  # :330-355
        "lrelu",
        [
            (32, 3, 2, False, True, "act", 0.0),
            (128, 3, 2, True, False, "act", 0.10),
            (128, 3, 2, False, True, "act", 0.0),
        ],
    ),
    "depth_ver_1": (  # :358-392
        "lrelu",
        [
            (16, 3, 2, False, True, "act", 0.0),
            (64, 3, 2, True, False, "act", 0.10),
            (128, 3, 1, True, False, "act", 0.15),
            (128, 3, 2, False, True, "act", 0.0),
            (128, 3, 1, True, False, "act", 0.0),
        ],
    ),
    "depth_ver_2": ("lrelu", _eight(16, 32, 64, 128)),  # :395-397
    "depth_ver_3": (  # :400-459
        "lrelu",
        [
            (16, 3, 2, False, True, "act", 0.0),
            (32, 3, 1, True, False, "act", 0.05),
            (32, 3, 1, True, False, "act", 0.05),
            (64, 3, 2, True, False, "act", 0.10),
            (128, 3, 1, True, False, "act", 0.15),
            (128, 3, 1, True, True, "act", 0.0),
            (128, 3, 2, False, False, "act", 0.0),
            (128, 3, 1, True, True, "act", 0.0),
            (128, 3, 1, True, False, "act", 0.0),
        ],
    ),
    "depth_ver_4": (  # :462-529
        "lrelu",
        [
            (16, 3, 2, False, True, "act", 0.0),
            (16, 3, 1, True, False, "act", 0.0),
            (32, 3, 1, True, False, "act", 0.05),
            (32, 3, 1, True, False, "act", 0.05),
            (64, 3, 2, True, False, "act", 0.10),
            (64, 3, 1, True, False, "act", 0.0),
            (128, 3, 1, True, False, "act", 0.15),
            (128, 3, 1, True, True, "act", 0.0),
            (128, 3, 2, True, False, "act", 0.0),
            (128, 3, 1, True, True, "act", 0.0),
            (128, 3, 1, True, False, "act", 0.0),
        ],
    ),
}


@dataclass
class Block:
    weight: torch.Tensor  # OIHW fp32
    bias: Optional[torch.Tensor]
    stride: int
    pad: int
    bn: Optional[Dict[str, torch.Tensor]]  # weight, bias, running_mean, running_var
    act: Optional[str]  # "lrelu" | "silu" | None
    p_drop: float


def blocks_from_state_dict(model_name: str, sd: Dict[str, torch.Tensor]) -> List[Block]:
    """Bind a reference-layout state_dict (SURVEY.md Appendix D key list) to the table."""
    act, rows = BACKBONES[model_name]
    blocks: List[Block] = []
    for i, (cout, k, s, has_bias, has_bn, a, p) in enumerate(rows):
        w = sd[f"model.{i}.0.weight"].detach().float().cpu()
        assert w.shape[0] == cout and w.shape[2] == k
        b = sd[f"model.{i}.0.bias"].detach().float().cpu() if has_bias else None
        bn = None
        if has_bn:
            bn = {
                kk: sd[f"model.{i}.1.{kk}"].detach().float().cpu().clone()
                for kk in ("weight", "bias", "running_mean", "running_var")
            }
        blocks.append(Block(w, b, s, k // 2, bn, act if a else None, p))
    i = len(rows)  # the bare 1x1 head conv, e.g. model_defns.py:67
    blocks.append(
        Block(
            sd[f"model.{i}.weight"].detach().float().cpu(),
            sd[f"model.{i}.bias"].detach().float().cpu(),
            1,
            0,
            None,
            None,
            0.0,
        )
    )
    return blocks


def _act(x: torch.Tensor, act: Optional[str]) -> torch.Tensor:
    if act == "lrelu":
        return F.leaky_relu(x, 0.01)  # nn.LeakyReLU() default slope (model_defns.py:36)
    if act == "silu":
        return F.silu(x)  # model_defns.py:86
    return x


def backbone_forward(
    x: torch.Tensor,
    blocks: Sequence[Block],
    train: bool,
    drop_keep: Optional[Sequence[Optional[torch.Tensor]]] = None,
    update_running: bool = False,
    dtype: torch.dtype = torch.float32,
) -> torch.Tensor:
    """nn.Sequential of conv blocks (model_defns.py:68-77): conv -> [BN] -> act -> [Dropout2d].

    drop_keep[i] is a (N, C) 0/1 keep mask for block i (Dropout2d zeroes whole (n, c)
    planes and scales the rest by 1/(1-p)); None = no dropout (eval, or p = 0).
    """
    x = x.to(dtype)
    for i, blk in enumerate(blocks):
        w = blk.weight.to(dtype)
        b = None if blk.bias is None else blk.bias.to(dtype)
        x = F.conv2d(x, w, b, stride=blk.stride, padding=blk.pad)
        if blk.bn is not None:
            rm, rv = blk.bn["running_mean"], blk.bn["running_var"]
            if not update_running:
                rm, rv = rm.clone(), rv.clone()
            x = F.batch_norm(
                x,
                rm.to(dtype) if not update_running else rm,
                rv.to(dtype) if not update_running else rv,
                blk.bn["weight"].to(dtype),
                blk.bn["bias"].to(dtype),
                training=train,
                momentum=0.1,
                eps=1e-5,
            )
        x = _act(x, blk.act)
        if drop_keep is not None and drop_keep[i] is not None and blk.p_drop > 0:
            keep = drop_keep[i].to(dtype)
            x = x * (keep / (1.0 - blk.p_drop))[:, :, None, None]
    return x


def head_transform(
    t: torch.Tensor,
    anchor_w: float,
    anchor_h: float,
    width_multiplier: float = 1.0,
    height_multiplier: float = 1.0,
    inference: bool = False,
) -> torch.Tensor:
    """YOGO.forward after the backbone (/root/reference/yogo/model.py:277-313)."""
    _, _, Sy, Sx = t.shape
    # model.py:48-55 - linspace(0, 1 - 1/S, S)
    cx = torch.linspace(0, 1 - 1 / Sx, Sx, dtype=torch.float32).to(t.dtype).expand(Sy, -1)
    cy = torch.linspace(0, 1 - 1 / Sy, Sy, dtype=torch.float32).to(t.dtype)[:, None].expand(Sy, Sx)
    cls = torch.softmax(t[:, 5:], dim=1) if inference else t[:, 5:]
    wh = torch.clamp(t[:, 2:4], max=80)
    return torch.cat(
        (
            ((1 / Sx) * torch.sigmoid(t[:, 0]) + cx)[:, None],
            ((1 / Sy) * torch.sigmoid(t[:, 1]) + cy)[:, None],
            anchor_w * torch.exp(wh[:, 0:1]) * width_multiplier,
            anchor_h * torch.exp(wh[:, 1:2]) * height_multiplier,
            torch.sigmoid(t[:, 4])[:, None],
            cls,
        ),
        dim=1,
    )


# ----------------------------------------------------------------------------------------
# YOGOLoss (/root/reference/yogo/yogo_loss.py:38-129; torchvision ciou_loss.py/diou_loss.py)
# ----------------------------------------------------------------------------------------
def yogo_loss_np(
    pred: np.ndarray,
    label: np.ndarray,
    no_obj_weight: float = 0.5,
    iou_weight: float = 5.0,
    classify_weight: float = 1.0,
    label_smoothing: float = 0.01,
    want_grad: bool = True,
    dtype=np.float32,
):
    """Per-cell restatement.  Returns (loss, {"iou_loss","objectness_loss","classification_loss"}, dpred).

    pred (N, 5+C, Sy, Sx), label (N, 6, Sy, Sx) = [mask, x1, y1, x2, y2, cls].
    Sums are accumulated in float64; per-cell arithmetic is done in `dtype`.
    """
    pred = np.asarray(pred, dtype=dtype)
    label = np.asarray(label, dtype=dtype)
    N, D, Sy, Sx = pred.shape
    C = D - 5
    f = dtype
    eps = f(1e-7)
    dpred = np.zeros_like(pred)

    m = label[:, 0]
    mb = m != 0  # .bool() (yogo_loss.py:69-73)

    # ---- objectness (yogo_loss.py:116-119) ----
    wobj = m * f(1 - no_obj_weight) + f(no_obj_weight)
    d = pred[:, 4] - m
    obj_sum = float(np.sum((d * d * wobj).astype(np.float64)))
    if want_grad:
        dpred[:, 4] = f(2.0) * d * wobj / f(N)

    # ---- classification (yogo_loss.py:107-114); CrossEntropyLoss(label_smoothing) ----
    logits = pred[:, 5:]
    mx = logits.max(axis=1, keepdims=True)
    ex = np.exp(logits - mx)
    se = ex.sum(axis=1, keepdims=True)
    logp = logits - mx - np.log(se)
    tgt = label[:, 5].astype(np.int64)
    tgt_c = np.clip(tgt, 0, C - 1)
    nll = -np.take_along_axis(logp, tgt_c[:, None], axis=1)[:, 0]
    smooth = -logp.mean(axis=1)
    ce = f(1 - label_smoothing) * nll + f(label_smoothing) * smooth
    cls_sum = float(np.sum((m * ce).astype(np.float64)))
    if want_grad:
        sm = ex / se
        onehot = np.zeros_like(sm)
        np.put_along_axis(onehot, tgt_c[:, None], 1.0, axis=1)
        tdist = f(1 - label_smoothing) * onehot + f(label_smoothing / C)
        dpred[:, 5:] = (m[:, None] * (sm - tdist)) * f(classify_weight) / f(N)

    # ---- box term (yogo_loss.py:59-105) ----
    idx = np.nonzero(mb)
    p0, p1, p2, p3 = (pred[:, k][idx] for k in range(4))
    gx1, gy1, gx2, gy2 = (label[:, k][idx] for k in range(1, 5))
    # torchvision _box_cxcywh_to_xyxy
    bx1 = p0 - f(0.5) * p2
    by1 = p1 - f(0.5) * p3
    bx2 = p0 + f(0.5) * p2
    by2 = p1 + f(0.5) * p3
    valid = (bx1 != bx2) & (by1 != by2)  # tested before clamping (yogo_loss.py:84-90)
    # clamp to [0,1] (yogo_loss.py:96-100); gradient passes where 0 <= v <= 1
    pas = [((v >= 0) & (v <= 1)) for v in (bx1, by1, bx2, by2)]
    x1, y1, x2, y2 = (np.clip(v, f(0), f(1)) for v in (bx1, by1, bx2, by2))
    with np.errstate(divide="ignore", invalid="ignore"):
        xk1 = np.maximum(x1, gx1)
        yk1 = np.maximum(y1, gy1)
        xk2 = np.minimum(x2, gx2)
        yk2 = np.minimum(y2, gy2)
        imask = (yk2 > yk1) & (xk2 > xk1)
        inter = np.where(imask, (xk2 - xk1) * (yk2 - yk1), f(0))
        union = (x2 - x1) * (y2 - y1) + (gx2 - gx1) * (gy2 - gy1) - inter
        iou = inter / (union + eps)
        xc1 = np.minimum(x1, gx1)
        yc1 = np.minimum(y1, gy1)
        xc2 = np.maximum(x2, gx2)
        yc2 = np.maximum(y2, gy2)
        diag = (xc2 - xc1) ** 2 + (yc2 - yc1) ** 2 + eps
        xp = (x2 + x1) / f(2)
        yp = (y2 + y1) / f(2)
        xg = (gx1 + gx2) / f(2)
        yg = (gy1 + gy2) / f(2)
        cent = (xp - xg) ** 2 + (yp - yg) ** 2
        wp = x2 - x1
        hp = y2 - y1
        wg = gx2 - gx1
        hg = gy2 - gy1
        k = f(4 / (math.pi**2))
        dat = np.arctan(wg / hg) - np.arctan(wp / hp)
        v = k * dat * dat
        alpha = v / (f(1) - iou + v + eps)  # no_grad
        ciou = f(1) - iou + cent / diag + alpha * v
    iou_sum = float(np.sum(np.where(valid, ciou, f(0)).astype(np.float64)))

    if want_grad and len(p0):
        with np.errstate(divide="ignore", invalid="ignore"):

            def sel_max(a, b):  # d max(a,b)/da; ties split (torch maximum backward)
                return np.where(a > b, f(1), np.where(a == b, f(0.5), f(0)))

            def sel_min(a, b):
                return np.where(a < b, f(1), np.where(a == b, f(0.5), f(0)))

            wI = xk2 - xk1
            hI = yk2 - yk1
            im = imask.astype(f)
            dI = [
                -hI * sel_max(x1, gx1) * im,
                -wI * sel_max(y1, gy1) * im,
                hI * sel_min(x2, gx2) * im,
                wI * sel_min(y2, gy2) * im,
            ]
            dA = [-hp, -wp, hp, wp]  # d area_pred
            dU = [dA[i] - dI[i] for i in range(4)]
            Ue = union + eps
            diou = [(dI[i] * Ue - inter * dU[i]) / (Ue * Ue) for i in range(4)]
            dD = [
                -f(2) * (xc2 - xc1) * sel_min(x1, gx1),
                -f(2) * (yc2 - yc1) * sel_min(y1, gy1),
                f(2) * (xc2 - xc1) * sel_max(x2, gx2),
                f(2) * (yc2 - yc1) * sel_max(y2, gy2),
            ]
            dC = [xp - xg, yp - yg, xp - xg, yp - yg]
            r = wp / hp
            dv_dw = k * f(2) * (-dat) * (f(1) / (f(1) + r * r)) / hp
            dv_dh = k * f(2) * (-dat) * (f(1) / (f(1) + r * r)) * (-wp / (hp * hp))
            dv = [-dv_dw, -dv_dh, dv_dw, dv_dh]
            g = [
                (-diou[i] + (dC[i] * diag - cent * dD[i]) / (diag * diag) + alpha * dv[i])
                for i in range(4)
            ]
            scale = f(iou_weight) / f(N)
            g = [np.where(valid & pas[i], g[i], f(0)) * scale for i in range(4)]
        # chain through cxcywh -> xyxy
        d0 = g[0] + g[2]
        d1 = g[1] + g[3]
        d2 = f(0.5) * (g[2] - g[0])
        d3 = f(0.5) * (g[3] - g[1])
        for kk, dk in enumerate((d0, d1, d2, d3)):
            plane = dpred[:, kk]
            plane[idx] = dk

    iou_loss = iou_weight * iou_sum / N
    cls_loss = classify_weight * cls_sum / N
    obj_loss = obj_sum / N
    comps = {
        "iou_loss": iou_loss,
        "objectness_loss": obj_loss,
        "classification_loss": cls_loss,
    }
    return obj_loss + iou_loss + cls_loss, comps, (dpred if want_grad else None)


# ----------------------------------------------------------------------------------------
# format_preds + torchvision.ops.nms (CPU semantics) + class counts
# (/root/reference/yogo/utils/prediction_formatting.py:23-93, yogo/infer.py:60-124)
# ----------------------------------------------------------------------------------------
def nms_np(boxes: np.ndarray, scores: np.ndarray, iou_threshold: float) -> np.ndarray:
    """Greedy NMS with torchvision CPU semantics: stable descending
    sort, fp32 arithmetic without fusion, strict '>' against the threshold as a double,
    NaN never suppresses.  Returns kept indices (int64) in descending-score order."""
    boxes = np.asarray(boxes, dtype=np.float32)
    scores = np.asarray(scores, dtype=np.float32)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    order = np.argsort(-scores, kind="stable")
    x1, y1, x2, y2 = (boxes[order, i] for i in range(4))
    areas = (x2 - x1) * (y2 - y1)
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    zero = np.float32(0)
    thr = float(iou_threshold)
    for i in range(n):
        if suppressed[i]:
            continue
        keep.append(order[i])
        if i + 1 == n:
            break
        xx1 = np.maximum(x1[i], x1[i + 1 :])
        yy1 = np.maximum(y1[i], y1[i + 1 :])
        xx2 = np.minimum(x2[i], x2[i + 1 :])
        yy2 = np.minimum(y2[i], y2[i + 1 :])
        w = np.maximum(zero, xx2 - xx1)
        h = np.maximum(zero, yy2 - yy1)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / ((areas[i] + areas[i + 1 :]) - inter)
        suppressed[i + 1 :] |= ovr.astype(np.float64) > thr
    return np.asarray(keep, dtype=np.int64)


def format_preds_np(
    pred: np.ndarray,
    obj_thresh: float = 0.5,
    iou_thresh: float = 0.5,
    box_format: str = "cxcywh",
    min_class_confidence_threshold: float = 0.0,
    return_index: bool = False,
):
    if pred.ndim != 3:
        raise ValueError(
            "argument to format_pred should be unbatched result - "
            f"shape should be (pred_shape, Sy, Sx), got {pred.shape}"
        )
    if box_format not in ("xyxy", "cxcywh"):
        raise ValueError(f"invalid box format {box_format}")
    D, Sy, Sx = pred.shape
    rows = np.asarray(pred, dtype=np.float32).reshape(D, Sy * Sx).T  # row k = cell j*Sx+i
    cand = np.nonzero(rows[:, 4] > np.float32(obj_thresh))[0]
    p = rows[cand].copy()
    half = np.float32(0.5)
    xyxy = np.stack(
        (
            p[:, 0] - half * p[:, 2],
            p[:, 1] - half * p[:, 3],
            p[:, 0] + half * p[:, 2],
            p[:, 1] + half * p[:, 3],
        ),
        axis=1,
    ).astype(np.float32)
    if box_format == "xyxy":
        p[:, :4] = xyxy
    index = cand
    if iou_thresh > 0:
        score = p[:, 5:].max(axis=1) * p[:, 4] if len(p) else np.zeros((0,), np.float32)
        keep = nms_np(xyxy, score, iou_thresh)
        p = p[keep]
        index = index[keep]
    if min_class_confidence_threshold > 0:
        k2 = p[:, 5:].max(axis=1) > np.float32(min_class_confidence_threshold)
        p = p[k2]
        index = index[k2]
    return (p, index) if return_index else p


def count_cells_np(cls_rows: np.ndarray, min_confidence_threshold: Optional[float] = None) -> np.ndarray:
    """infer.py:90-124: argmax (first max wins) histogram of rows whose max > threshold."""
    if cls_rows.ndim != 2:
        raise ValueError("expected formatted_class_predictions to be shape (N, num_classes)")
    if min_confidence_threshold is not None:
        if min_confidence_threshold < 0 or min_confidence_threshold > 1:
            raise ValueError("min_confidence_threshold should be between 0 and 1")
    else:
        min_confidence_threshold = 0
    C = cls_rows.shape[1]
    if cls_rows.shape[0] == 0:
        return np.zeros(C, dtype=np.int64)
    vals = cls_rows.max(axis=1)
    arg = cls_rows.argmax(axis=1)
    sel = arg[vals > np.float32(min_confidence_threshold)]
    return np.bincount(sel, minlength=C).astype(np.int64)


def prediction_class_counts_np(
    batch_preds: np.ndarray,
    obj_thresh: float = 0.5,
    iou_thresh: float = 0.5,
    min_class_confidence_threshold: float = 0.0,
) -> np.ndarray:
    """infer.py:60-87."""
    C = batch_preds.shape[1] - 5
    tot = np.zeros(C, dtype=np.int64)
    for sl in batch_preds:
        p = format_preds_np(sl, obj_thresh, iou_thresh, "cxcywh", min_class_confidence_threshold)
        if p.size == 0:
            continue
        tot += count_cells_np(p[:, 5:])
    return tot


# ----------------------------------------------------------------------------------------
# Evaluation matching: format_preds_and_labels_v2
# (/root/reference/yogo/utils/prediction_formatting.py:254-330; torchvision.ops.box_iou, boxes.py _box_inter_union)
# ----------------------------------------------------------------------------------------
def box_iou_cost_np(labels_xyxy: np.ndarray, preds_xyxy: np.ndarray) -> np.ndarray:
    """1 - box_iou(labels, preds) in fp32, operation by operation as torchvision computes it."""
    a = np.asarray(labels_xyxy, dtype=np.float32).reshape(-1, 4)
    b = np.asarray(preds_xyxy, dtype=np.float32).reshape(-1, 4)
    area1 = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area2 = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = np.maximum(a[:, None, :2], b[None, :, :2])
    rb = np.minimum(a[:, None, 2:], b[None, :, 2:])
    wh = rb - lt
    wh = np.where(wh < 0, np.float32(0), wh)
    inter = wh[:, :, 0] * wh[:, :, 1]
    union = (area1[:, None] + area2[None, :]) - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = inter / union
    return (np.float32(1) - iou).astype(np.float32)


def match_preds_and_labels_np(
    pred: np.ndarray, label: np.ndarray, objectness_thresh: float = 0.5, min_class_confidence_threshold: float = 0.0
):
    """Returns (matched preds, matched labels, missed labels, extra predictions) like PredictionLabelMatch."""
    from scipy.optimize import linear_sum_assignment

    fp = format_preds_np(pred, objectness_thresh, 0.5, "xyxy", min_class_confidence_threshold)
    Ls, Sy, Sx = label.shape
    labels = np.asarray(label, dtype=np.float32).reshape(Ls, Sy * Sx).T
    fl = labels[labels[:, 0] != 0]
    cost = box_iou_cost_np(fl[:, 1:5], fp[:, :4])
    rows, cols = linear_sum_assignment(cost)
    pred_free = np.ones(fp.shape[0], dtype=bool)
    pred_free[cols] = False
    label_free = np.ones(fl.shape[0], dtype=bool)
    label_free[rows] = False
    return fp[cols], fl[rows], fl[label_free], fp[pred_free]


def synth_labels_for_preds(pred: torch.Tensor, drop: float = 0.2, extra: int = 5, seed: int = 3) -> torch.Tensor:
    """Label tensor (6, Sy, Sx) for one synthetic prediction tensor (5 + C, Sy, Sx): the boxes that survive
    threshold + NMS, jittered by a few percent, minus a fraction `drop` (those predictions become extras), plus `extra`
    anchor-sized labels at random empty cells (missed labels).  Layout [mask, x1, y1, x2, y2, class]."""
    g = torch.Generator().manual_seed(seed)
    D, Sy, Sx = pred.shape
    rows = format_preds_np(pred.numpy(), 0.5, 0.5, "xyxy", 0.0)
    lab = torch.zeros(6, Sy, Sx)
    for r in rows:
        if float(torch.rand((), generator=g)) < drop:
            continue
        jit = (torch.rand(4, generator=g) - 0.5) * 0.004
        box = torch.from_numpy(r[:4].copy()) + jit
        cx, cy = float((box[0] + box[2]) / 2), float((box[1] + box[3]) / 2)
        i, j = min(max(int(cx * Sx), 0), Sx - 1), min(max(int(cy * Sy), 0), Sy - 1)
        lab[0, j, i] = 1
        lab[1:5, j, i] = box
        lab[5, j, i] = float(int(np.argmax(r[5:])))
    for _ in range(extra):
        j = int(torch.randint(0, Sy, (), generator=g))
        i = int(torch.randint(0, Sx, (), generator=g))
        if lab[0, j, i] != 0:
            continue
        cx, cy = (i + 0.5) / Sx, (j + 0.5) / Sy
        lab[0, j, i] = 1
        lab[1:5, j, i] = torch.tensor([cx - ANCHOR_W / 2, cy - ANCHOR_H / 2, cx + ANCHOR_W / 2, cy + ANCHOR_H / 2])
        lab[5, j, i] = float(int(torch.randint(0, D - 5, (), generator=g)))
    return lab


# ----------------------------------------------------------------------------------------
# Input side: bounding-box aware flips and label rasterisation
# (/root/reference/yogo/data/data_transforms.py:51-98, /root/reference/yogo/data/yogo_dataset.py:24-46)
# ----------------------------------------------------------------------------------------
def flip_batch_np(img: np.ndarray, lab: np.ndarray, hflip: bool, vflip: bool):
    img = np.asarray(img)
    lab = np.asarray(lab, dtype=np.float32).copy()
    one = np.float32(1)
    if hflip:
        x1, x2 = one - lab[:, 3], one - lab[:, 1]
        lab[:, 1], lab[:, 3] = x1, x2
        img, lab = img[..., ::-1], lab[..., ::-1]
    if vflip:
        y1, y2 = one - lab[:, 4], one - lab[:, 2]
        lab[:, 2], lab[:, 4] = y1, y2
        img, lab = img[..., ::-1, :], lab[..., ::-1, :]
    return np.ascontiguousarray(img), np.ascontiguousarray(lab)


def format_labels_tensor_np(labels: np.ndarray, Sx: int, Sy: int) -> np.ndarray:
    """(n, 5) [class, x1, y1, x2, y2] -> (6, Sy, Sx) [mask, x1, y1, x2, y2, class]; later labels overwrite earlier ones."""
    labels = np.asarray(labels, dtype=np.float32).reshape(-1, 5)
    out = np.zeros((6, Sy, Sx), dtype=np.float32)
    for lab in labels:
        i = int(np.floor(((lab[1] + lab[3]) * np.float32(Sx)) / np.float32(2)))
        j = int(np.floor(((lab[2] + lab[4]) * np.float32(Sy)) / np.float32(2)))
        if not (-Sx <= i < Sx and -Sy <= j < Sy):
            raise IndexError("label centre outside the grid")
        out[0, j, i] = 1
        out[1:5, j, i] = lab[1:]
        out[5, j, i] = lab[0]
    return out


# ----------------------------------------------------------------------------------------
# Synthetic inputs live in tools/synth.py (shared with bench.py); re-exported for the tests
# ----------------------------------------------------------------------------------------
from tools.synth import ANCHOR_H, ANCHOR_W, synth_images, synth_labels, synth_sparse_preds  # noqa: E402,F401
