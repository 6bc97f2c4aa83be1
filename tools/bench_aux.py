"""HBM-bound side kernels of the path against the copy-bandwidth roofline (CUDA events, inputs larger than L2 or L2 flushed):
the bounding-box aware flips and label rasterisation, the matching cost matrix (N2) and YOGOLoss.

    python tools/bench_aux.py [--batch 256] [--reps 20]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import yogo_b200  # noqa: E402
from yogo_b200.data import flip_batch, format_labels_batch  # noqa: E402
from yogo_b200.utils import box_iou_cost  # noqa: E402
from tools import synth as S  # noqa: E402


def timed(fn, reps, flush=None):
    fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        if flush is not None:
            flush.add_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    peak = float(peaks["hbm_gbs"])
    flush = torch.zeros(160 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)   # 160 MB > 126 MB L2
    B = args.batch
    img = S.synth_images(B).to(dev)
    lab = S.synth_labels(B).to(dev)

    def rec(name, ms, nbytes, unit_count, unit, cpu=None):
        r = {"kernel": name, "ms": round(ms, 4), "algorithmic_bytes": int(nbytes), "GBps": round(nbytes / ms / 1e6, 1),
             "frac_hbm_peak": round(nbytes / ms / 1e6 / peak, 3), unit + "_per_s": round(unit_count / ms * 1e3, 1)}
        if cpu is not None:
            r["cpu_reference_" + unit + "_per_s"] = round(cpu, 1)
        print(json.dumps(r), flush=True)

    # flips: image read + write, label read + write
    for tag, h, v in (("hflip", True, False), ("vflip", False, True), ("hvflip", True, True)):
        ms = timed(lambda: flip_batch(img, lab, h, v), args.reps)
        cpu = None
        if tag == "hflip":
            ci, cl = img[:16].cpu(), lab[:16].cpu()
            t0 = time.perf_counter()
            for _ in range(3):
                l2 = cl.clone()
                l2[:, 1], l2[:, 3] = 1 - l2[:, 3], 1 - l2[:, 1]
                torch.flip(ci, dims=(3,)), torch.flip(l2, dims=(3,))
            cpu = 16 * 3 / (time.perf_counter() - t0)
        rec("flip_images+labels[" + tag + "]", ms, 2 * img.numel() + 2 * lab.numel() * 4, B, "img", cpu)
    from yogo_b200 import _lib as L
    lib = L.lib()
    io, lo = torch.empty_like(img), torch.empty_like(lab)
    ms = timed(lambda: L.check(lib.yg_flip_images(img.data_ptr(), io.data_ptr(), L.YG_U8, B, 1, 772, 1032, 1, 1, L.stream())), args.reps)
    rec("yg_flip_images[hv, uint8]", ms, 2 * img.numel(), B, "img")
    ms = timed(lambda: L.check(lib.yg_flip_labels(lab.data_ptr(), lo.data_ptr(), B, 97, 129, 1, 1, L.stream())), args.reps)
    rec("yg_flip_labels[hv]", ms, 2 * lab.numel() * 4, B, "img")
    # label rasterisation: 300 labels per image
    g = torch.Generator().manual_seed(0)
    lists = []
    for _ in range(B):
        c = torch.rand(300, 2, generator=g)
        lists.append(torch.cat([torch.randint(0, 7, (300, 1), generator=g).float(), c - 0.02, c + 0.02], 1).clamp(0, 0.999))
    dl = [t.to(dev) for t in lists]
    ms = timed(lambda: format_labels_batch(dl, 129, 97), args.reps, flush)
    from oracle import yogo_oracle as O   # CPU baseline only
    t0 = time.perf_counter()
    for t in lists[:16]:
        O.format_labels_tensor_np(t.numpy(), 129, 97)
    rec("format_labels_batch (incl. host offsets + concat)", ms, B * (6 * 97 * 129 * 4 + 97 * 129 * 4 * 2 + 300 * 20), B, "img",
        16 / (time.perf_counter() - t0))
    # matching cost matrix: 700 labels x 2347 predictions (sparse-realistic K = 1000 candidates)
    a = torch.rand(700, 4, device=dev)
    b = torch.rand(2347, 12, device=dev)
    ms = timed(lambda: box_iou_cost(a, b), args.reps, flush)
    rec("box_iou_cost[700x2347]", ms, 700 * 2347 * 4 + 700 * 16 + 2347 * 48, 1, "matrix")
    # YOGOLoss forward + gradient: the C-ABI call (two launches), 72 B read + 48 B written per cell
    pred = torch.rand(B, 12, 97, 129, device=dev)
    dpred = torch.empty_like(pred)
    out4 = torch.empty(4, device=dev)
    nb = lib.yg_yogo_loss_workspace(B, 97, 129)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    ms = timed(lambda: L.check(lib.yg_yogo_loss_fwd_bwd(pred.data_ptr(), lab.data_ptr(), out4.data_ptr(), dpred.data_ptr(), B, 7,
                                                        97, 129, 0.5, 5.0, 1.0, 0.01, ws.data_ptr(), nb, L.stream())), args.reps)
    rec("yg_yogo_loss_fwd_bwd", ms, B * 97 * 129 * 120, B, "img")


if __name__ == "__main__":
    main()
