"""``YOGOLoss`` - same interface as /root/reference/yogo/yogo_loss.py:8-129, computed by one
fused CUDA kernel (forward + gradient in the same pass, csrc/loss.cu)."""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from . import _lib as L


class _LazyFloats(dict):
    """``loss_components`` dict whose values are read from the device on first access.

    The reference calls ``.item()`` three times inside ``forward`` (yogo_loss.py:123-127),
    forcing a host sync every step; here the sync happens only if somebody reads a value."""

    def __init__(self, dev4: torch.Tensor):
        super().__init__()
        self._dev4 = dev4
        self._done = False

    def _fill(self):
        if not self._done:
            v = self._dev4.tolist()
            dict.__setitem__(self, "iou_loss", v[1])
            dict.__setitem__(self, "objectness_loss", v[2])
            dict.__setitem__(self, "classification_loss", v[3])
            self._done = True

    def __getitem__(self, k):
        self._fill()
        return dict.__getitem__(self, k)

    def __iter__(self):
        self._fill()
        return dict.__iter__(self)

    def __len__(self):
        return 3

    def keys(self):
        self._fill()
        return dict.keys(self)

    def items(self):
        self._fill()
        return dict.items(self)

    def values(self):
        self._fill()
        return dict.values(self)

    def get(self, k, default=None):
        self._fill()
        return dict.get(self, k, default)

    def __contains__(self, k):
        return k in ("iou_loss", "objectness_loss", "classification_loss")

    def __repr__(self):
        self._fill()
        return dict.__repr__(self)


@L.on_device(lambda pred, *a: pred)
def _launch(pred, label, weights, need_grad):
    """One launch of the fused kernel: (out4 = [loss, iou, objectness, classification], d(loss)/d(pred) or None)."""
    lib = L.lib()
    N, D, Sy, Sx = pred.shape
    pred_c = pred.detach().contiguous().float()
    label_c = label.detach().contiguous().float()
    out4 = torch.empty(4, dtype=torch.float32, device=pred.device)
    dpred = torch.empty_like(pred_c) if need_grad else None
    nbytes = lib.yg_yogo_loss_workspace(N, Sy, Sx)
    ws = L.workspace.get("loss", nbytes, pred.device)
    no_obj, iou_w, cls_w, smooth = weights
    L.check(lib.yg_yogo_loss_fwd_bwd(pred_c.data_ptr(), label_c.data_ptr(), out4.data_ptr(), L.ptr(dpred),
                                     N, D - 5, Sy, Sx, no_obj, iou_w, cls_w, smooth, ws.data_ptr(), nbytes,
                                     L.stream()))
    return out4, dpred


class _LossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, label, weights, out4_holder):
        out4, dpred = _launch(pred, label, weights, pred.requires_grad)
        ctx.dpred = dpred
        out4_holder.append(out4)
        return out4[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        dpred = ctx.dpred
        ctx.dpred = None
        if dpred is None:
            return None, None, None, None
        return dpred * grad_out, None, None, None


class YOGOLoss(torch.nn.modules.loss._Loss):
    __constants__ = ["no_obj_weight", "iou_weight", "classify_weight"]

    def __init__(
        self,
        no_obj_weight: float = 0.5,
        iou_weight: float = 5.0,
        classify_weight: float = 1.0,
        label_smoothing: float = 0.01,
    ) -> None:
        super().__init__()
        self.no_obj_weight = no_obj_weight
        self.iou_weight = iou_weight
        self.classify_weight = classify_weight
        self.label_smoothing = label_smoothing
        self.device = "cpu"

    def to(self, device):
        self.device = device
        super().to(device, non_blocking=True, dtype=torch.float32)
        return self

    def forward(self, pred_batch: torch.Tensor, label_batch: torch.Tensor) -> Tuple[torch.Tensor, Dict[str, float]]:
        """pred (N, 5+C, Sy, Sx), label (N, 6, Sy, Sx) = [mask, x1, y1, x2, y2, class].
        Returns (loss, {"iou_loss", "objectness_loss", "classification_loss"})."""
        self._check(pred_batch, label_batch)
        holder: list = []
        loss = _LossFunction.apply(pred_batch, label_batch, self._weights(), holder)
        return loss, _LazyFloats(holder[0])

    def loss_and_grad(self, pred_batch: torch.Tensor, label_batch: torch.Tensor):
        """(loss, components, d(loss)/d(pred)) from the same single launch, without an autograd node.  The trainer hands
        the gradient straight to the network's backward: through autograd, `loss.backward()` first fills a tensor with the
        incoming gradient 1 and multiplies the 38 MB gradient by it (two extra launches per step)."""
        self._check(pred_batch, label_batch)
        out4, dpred = _launch(pred_batch, label_batch, self._weights(), True)
        return out4[0], _LazyFloats(out4), dpred

    def _weights(self):
        return (float(self.no_obj_weight), float(self.iou_weight), float(self.classify_weight), float(self.label_smoothing))

    @staticmethod
    def _check(pred_batch, label_batch) -> None:
        if pred_batch.ndim != 4 or label_batch.ndim != 4:
            raise ValueError("pred and label must be 4-d (N, C, Sy, Sx)")
        if pred_batch.shape[0] != label_batch.shape[0] or pred_batch.shape[2:] != label_batch.shape[2:]:
            raise RuntimeError(
                f"pred {tuple(pred_batch.shape)} and label {tuple(label_batch.shape)} do not describe the same grid")
        if label_batch.shape[1] != 6 or pred_batch.shape[1] < 6:
            raise RuntimeError("label must have 6 channels and pred 5 + num_classes")
        L.require_cuda(pred_batch, "YOGOLoss pred")
        L.require_cuda(label_batch, "YOGOLoss label")
