"""Batch assembly of the reference's DataLoader (yogo/data/utils.py:49-63)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from .data_transforms import MultiArgSequential


def collate_batch_robust(
    batch: List[Optional[Tuple[torch.Tensor, torch.Tensor]]],
    transforms: MultiArgSequential = MultiArgSequential(),
    device=None,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Drops the `None` items a dataset returned for unreadable files, stacks the rest and applies the batch transforms
    (yogo/data/utils.py:49-63: same signature and behaviour, including the ValueError of `zip(*[])` on an all-None
    batch).  `device` (an extension) moves the stacked batch there first, so that the flip transforms run as the CUDA
    kernels of csrc/input.cu on the whole batch; without it the tensors stay where the dataset put them."""
    inputs, labels = zip(*[pair for pair in batch if pair is not None])
    batched_inputs = torch.stack(inputs)
    batched_labels = torch.stack(labels)
    if device is not None:
        batched_inputs = batched_inputs.to(device, non_blocking=True)
        batched_labels = batched_labels.to(device, non_blocking=True)
    return transforms(batched_inputs, batched_labels)
