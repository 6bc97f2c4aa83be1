"""Batch transforms with the interface of /root/reference/yogo/data/data_transforms.py (applied by the collate function to
every training batch, yogo_dataloader.py:205-247).  The flips run as one image kernel and one label kernel per batch
(csrc/input.cu) on device-resident batches; the random decision is drawn exactly like the reference (``torch.rand(1) < p``
on the host generator), so a seeded run flips the same batches."""
from __future__ import annotations

from typing import Tuple

import torch

from .. import _lib as L


class DualInputModule(torch.nn.Module):
    def forward(self, inpt_a, inpt_b): ...


class DualInputId(DualInputModule):
    def forward(self, img_batch, labels):
        return img_batch, labels


class MultiArgSequential(torch.nn.Sequential):
    """data_transforms.py:26-35: identity transforms are dropped, every module maps (images, labels) -> (images, labels)."""

    def __init__(self, *args: DualInputModule, **kwargs):
        super().__init__(*[t for t in args if not isinstance(t, DualInputId)], **kwargs)

    def forward(self, *input):
        for module in self:
            input = module(*input)
        return input


class ImageTransformLabelIdentity(DualInputModule):
    """data_transforms.py:37-48: an image-only transform next to untouched labels."""

    def __init__(self, transform):
        super().__init__()
        self.transform = transform

    def forward(self, img_batch, labels):
        return self.transform(img_batch), labels


@L.on_device(lambda img_batch, *a, **k: img_batch)
def flip_batch(img_batch: torch.Tensor, label_batch: torch.Tensor, hflip: bool, vflip: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """Images (N,C,H,W) uint8/fp32 and labels (N,6,Sy,Sx) flipped horizontally and/or vertically with the box coordinates
    mirrored (x1' = 1 - x2, ...), out of place, two launches."""
    assert img_batch.ndim == 4 and label_batch.ndim == 4
    if not (hflip or vflip):
        return img_batch, label_batch
    L.require_cuda(img_batch, "images")
    L.require_cuda(label_batch, "labels")
    if img_batch.dtype not in (torch.uint8, torch.float32):
        raise TypeError(f"flip_batch: images must be uint8 or float32, got {img_batch.dtype}")
    if label_batch.shape[1] != 6:
        raise ValueError(f"labels must have shape (N, 6, Sy, Sx), got {tuple(label_batch.shape)}")
    lib = L.lib()
    img = img_batch.contiguous()
    lab = label_batch.contiguous().float()
    img_out = torch.empty_like(img)
    lab_out = torch.empty_like(lab)
    N, C, H, W = img.shape
    L.check(lib.yg_flip_images(img.data_ptr(), img_out.data_ptr(), L.dtype_code(img.dtype), N, C, H, W, int(hflip), int(vflip),
                               L.stream()))
    L.check(lib.yg_flip_labels(lab.data_ptr(), lab_out.data_ptr(), lab.shape[0], lab.shape[2], lab.shape[3], int(hflip),
                               int(vflip), L.stream()))
    return img_out, lab_out


class RandomHorizontalFlipWithBBs(DualInputModule):
    """data_transforms.py:51-73."""

    def __init__(self, p=0.5):
        super().__init__()
        self.p = p

    def forward(self, img_batch: torch.Tensor, label_batch: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        assert img_batch.ndim == 4 and label_batch.ndim == 4
        if torch.rand(1) < self.p:
            return flip_batch(img_batch, label_batch, True, False)
        return img_batch, label_batch


class RandomVerticalFlipWithBBs(DualInputModule):
    """data_transforms.py:76-98."""

    def __init__(self, p=0.5):
        super().__init__()
        self.p = p

    def forward(self, img_batch: torch.Tensor, label_batch: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        assert img_batch.ndim == 4 and label_batch.ndim == 4
        if torch.rand(1) < self.p:
            return flip_batch(img_batch, label_batch, False, True)
        return img_batch, label_batch
