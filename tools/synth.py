"""Seeded synthetic inputs of SURVEY.md 8d (images, label tensors, sparse-realistic prediction tensors).

Shared by bench.py, tools/, the tests and the oracle so that every arm sees the same data.  Input generation only:
nothing here computes any part of the path.
"""
import torch

ANCHOR_W = 0.04250100424705710  # default_hyperparams.py:12
ANCHOR_H = 0.05551774140353888  # default_hyperparams.py:11


def synth_images(B: int, H: int = 772, W: int = 1032, seed: int = 0, channels: int = 1) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (B, channels, H, W), dtype=torch.uint8, generator=g)


def synth_labels(B: int, Sy: int = 97, Sx: int = 129, C: int = 7, K: int = 300, seed: int = 1) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    lab = torch.zeros(B, 6, Sy, Sx, dtype=torch.float32)
    K = min(K, Sy * Sx)
    for b in range(B):
        cells = torch.randperm(Sy * Sx, generator=g)[:K]
        j = cells // Sx
        i = cells % Sx
        cx = (i + torch.rand(K, generator=g)) / Sx
        cy = (j + torch.rand(K, generator=g)) / Sy
        w = ANCHOR_W * (0.75 + 0.5 * torch.rand(K, generator=g))
        h = ANCHOR_H * (0.75 + 0.5 * torch.rand(K, generator=g))
        cls = torch.randint(0, C, (K,), generator=g).float()
        lab[b, 0, j, i] = 1.0
        lab[b, 1, j, i] = cx - w / 2
        lab[b, 2, j, i] = cy - h / 2
        lab[b, 3, j, i] = cx + w / 2
        lab[b, 4, j, i] = cy + h / 2
        lab[b, 5, j, i] = cls
    return lab


def synth_sparse_preds(B: int, Sy: int = 97, Sx: int = 129, C: int = 7, K: int = 300, seed: int = 2) -> torch.Tensor:
    """'sparse-realistic' prediction tensors: background cells with low
    objectness plus K objects per image lighting 1-4 cells of a 2x2 neighbourhood."""
    g = torch.Generator().manual_seed(seed)
    p = torch.zeros(B, 5 + C, Sy, Sx, dtype=torch.float32)
    jj, ii = torch.meshgrid(torch.arange(Sy), torch.arange(Sx), indexing="ij")
    p[:, 0] = (ii + 0.5) / Sx
    p[:, 1] = (jj + 0.5) / Sy
    p[:, 2] = ANCHOR_W
    p[:, 3] = ANCHOR_H
    p[:, 4] = 0.3 * torch.rand(B, Sy, Sx, generator=g)
    p[:, 5:] = torch.softmax(3 * torch.randn(B, C, Sy, Sx, generator=g), dim=1)
    K = min(K, (Sy - 1) * (Sx - 1))
    for b in range(B):
        cells = torch.randperm((Sy - 1) * (Sx - 1), generator=g)[:K]
        j0 = cells // (Sx - 1)
        i0 = cells % (Sx - 1)
        cx = (i0 + 1.0) / Sx
        cy = (j0 + 1.0) / Sy
        for dj in (0, 1):
            for di in (0, 1):
                lit = torch.rand(K, generator=g) < 0.625
                if dj == 0 and di == 0:
                    lit[:] = True
                j = (j0 + dj)[lit]
                i = (i0 + di)[lit]
                n = int(lit.sum())
                p[b, 0, j, i] = cx[lit] + 0.01 * ANCHOR_W * (2 * torch.rand(n, generator=g) - 1)
                p[b, 1, j, i] = cy[lit] + 0.01 * ANCHOR_H * (2 * torch.rand(n, generator=g) - 1)
                p[b, 2, j, i] = ANCHOR_W * (1 + 0.05 * (2 * torch.rand(n, generator=g) - 1))
                p[b, 3, j, i] = ANCHOR_H * (1 + 0.05 * (2 * torch.rand(n, generator=g) - 1))
                p[b, 4, j, i] = 0.6 + 0.4 * torch.rand(n, generator=g)
    return p


def synth_fill_model(net: torch.nn.Module, seed: int = 0, head_scale: float = 0.05) -> None:
    """Deterministic parameters for parity fixtures without committing the weights: values are integers drawn with
    numpy's PCG64 scaled by a power-of-two-free constant in fp32 (bit-reproducible on any host).  Same statistics as
    the reference's Kaiming(fan_out) init (model.py:79-87) for conv weights, head weights scaled by ``head_scale``
    (SURVEY.md 8d "tame"), conv biases U(-.1,.1), BatchNorm weight U(.5,1.5) / bias U(-.2,.2).  Applies to any module
    whose ``.model`` is the backbone Sequential (the reference's YOGO and yogo_b200.YOGO alike)."""
    import numpy as np

    rng = np.random.default_rng(seed)

    def u(shape, lo, hi):
        ints = rng.integers(0, 1 << 16, size=tuple(shape), dtype=np.int64).astype(np.float32)
        return torch.from_numpy(np.float32(lo) + ints * np.float32((hi - lo) / 65536.0))

    convs = [m for m in net.model.modules() if isinstance(m, torch.nn.Conv2d)]
    with torch.no_grad():
        for m in net.model.modules():
            if isinstance(m, torch.nn.Conv2d):
                fan_out = m.weight.shape[0] * m.weight.shape[2] * m.weight.shape[3]
                a = (3.0 * 2.0 / fan_out) ** 0.5
                if m is convs[-1]:
                    a *= head_scale
                m.weight.copy_(u(m.weight.shape, -a, a))
                if m.bias is not None:
                    m.bias.copy_(u(m.bias.shape, -0.1, 0.1))
            elif isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_(u(m.weight.shape, 0.5, 1.5))
                m.bias.copy_(u(m.bias.shape, -0.2, 0.2))
