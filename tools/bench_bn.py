"""BatchNorm streaming kernels on the 128-channel layers of base_model at batch 64 (CUDA events, 205 MB tensors > L2).

    python tools/bench_bn.py [--reps 20] [--tc-options N]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from yogo_b200 import _lib as L  # noqa: E402

if os.environ.get("YOGO_B200_LIB_ALT"):
    L.LIB_PATH = os.environ["YOGO_B200_LIB_ALT"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--tc-options", type=int, default=None)
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args()
    lib = L.lib()
    if args.tc_options is not None:
        lib.yg_set_tc_options(args.tc_options)
    dev = "cuda:0"
    N, HW, C = args.batch, 97 * 129, 128
    y = torch.randn(N, HW, C, device=dev).bfloat16()
    g = torch.randn(N, HW, C, device=dev).bfloat16()
    a = torch.empty_like(y)
    st = L.stream()
    mean = torch.zeros(C, device=dev)
    invstd = torch.ones(C, device=dev)
    scale = torch.ones(C, device=dev)
    shift = torch.zeros(C, device=dev)
    gamma = torch.ones(C, device=dev)
    mask = torch.empty(y.numel() // 8, dtype=torch.uint8, device=dev)
    nbytes = y.numel() * 2

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.reps

    stats = torch.zeros(2 * C, dtype=torch.float64, device=dev)
    sums = torch.zeros(2 * C, dtype=torch.float64, device=dev)
    cases = [
        ("yg_bn_stats", 1, lambda: L.check(lib.yg_bn_stats(y.data_ptr(), L.YG_BF16, N, HW, C, stats.data_ptr(), st))),
        ("yg_bn_act_apply", 2, lambda: L.check(lib.yg_bn_act_apply(y.data_ptr(), a.data_ptr(), L.YG_BF16, N, HW, C, scale.data_ptr(), shift.data_ptr(),
                                                                  L.ACT_LRELU, None, mask.data_ptr(), st))),
        ("yg_bn_bwd_sums", 2, lambda: L.check(lib.yg_bn_bwd_sums(g.data_ptr(), y.data_ptr(), L.YG_BF16, N, HW, C, mean.data_ptr(), invstd.data_ptr(),
                                                                sums.data_ptr(), st))),
        ("yg_bn_bwd_apply", 3, lambda: L.check(lib.yg_bn_bwd_apply(g.data_ptr(), y.data_ptr(), L.YG_BF16, N, HW, C, sums.data_ptr(), gamma.data_ptr(),
                                                                  mean.data_ptr(), invstd.data_ptr(), None, None, 1.0, 1, st))),
    ]
    for name, passes, fn in cases:
        ms = timed(fn)
        print(json.dumps({"kernel": name, "opt": args.tc_options, "ms": round(ms, 4), "GBps": round(passes * nbytes / ms / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    main()
