"""Label rasterisation of /root/reference/yogo/data/yogo_dataset.py:15-46: ragged (n, 5) label lists
``[class, x1, y1, x2, y2]`` -> the (6, Sy, Sx) tensor ``[mask, x1, y1, x2, y2, class]`` YOGOLoss consumes.
The reference loops over labels in Python per image inside every DataLoader worker; ``format_labels_batch`` rasterises a
whole batch on the GPU (csrc/input.cu) with the same cell arithmetic and the same last-label-wins rule."""
from __future__ import annotations

from typing import Sequence

import torch

from .. import _lib as L

LABEL_TENSOR_PRED_DIM_SIZE = 1 + 4 + 1


def format_labels_batch(labels: Sequence[torch.Tensor], Sx: int, Sy: int, device=None) -> torch.Tensor:
    """One (n_b, 5) tensor per image -> (B, 6, Sy, Sx) on the GPU.  Raises IndexError when a label centre falls outside the
    grid (the reference's ``output[0, j, i] = 1`` does)."""
    B = len(labels)
    dev = torch.device(device) if device is not None else (labels[0].device if B and labels[0].is_cuda else torch.device("cuda"))
    counts = [int(t.shape[0]) for t in labels]
    offs = torch.zeros(B + 1, dtype=torch.int32)
    if B:
        offs[1:] = torch.tensor(counts, dtype=torch.int64).cumsum(0).to(torch.int32)
    total = int(offs[-1])
    flat = (torch.cat([t.reshape(-1, 5).float() for t in labels]) if total else torch.zeros((0, 5))).to(dev).contiguous()
    offs_d = offs.to(dev)
    out = torch.empty((B, LABEL_TENSOR_PRED_DIM_SIZE, Sy, Sx), dtype=torch.float32, device=dev)
    owner = torch.empty((B, Sy, Sx), dtype=torch.int32, device=dev)
    err = torch.empty(1, dtype=torch.int32, device=dev)
    L.check(L.lib().yg_format_labels_batch(flat.data_ptr() if total else None, offs_d.data_ptr(), B, max(counts) if counts else 0,
                                           Sy, Sx, owner.data_ptr(), out.data_ptr(), err.data_ptr(), L.stream()))
    if B and int(err.item()):
        raise IndexError("a label centre falls outside the (Sy, Sx) grid")
    return out


def format_labels_tensor(labels: torch.Tensor, Sx: int, Sy: int) -> torch.Tensor:
    """yogo_dataset.py:24-46, one image: (N, 5) -> (6, Sy, Sx)."""
    return format_labels_batch([labels], Sx, Sy)[0]
