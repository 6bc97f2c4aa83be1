"""Diagnostic for the tcgen05 conv engine: runs each case in a subprocess (a device-side trap
kills the CUDA context) and reports where results differ from the SIMT kernels.

    python tools/tc_debug.py            # all cases
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    # kind, N, H, W, Cin, Cout, stride
    ("fwd", 1, 8, 16, 64, 64, 1),
    ("fwd", 1, 8, 16, 16, 32, 1),
    ("fwd", 1, 8, 16, 32, 64, 1),
    ("fwd", 2, 19, 35, 64, 128, 1),
    ("fwd", 2, 19, 35, 128, 128, 1),
    ("fwd", 2, 20, 36, 32, 64, 2),
    ("fwd", 2, 21, 37, 128, 128, 2),
    ("fwd", 1, 16, 32, 128, 256, 1),
    ("dgrad", 1, 8, 16, 64, 64, 1),
    ("dgrad", 2, 19, 35, 64, 128, 1),
    ("dgrad", 2, 20, 36, 32, 64, 2),
    ("dgrad", 2, 21, 37, 128, 128, 2),
    ("dgrad", 2, 21, 36, 32, 64, 2),
    ("dgrad", 2, 20, 37, 32, 64, 2),
    ("dgrad", 3, 37, 70, 16, 32, 2),
    ("dgrad", 2, 19, 35, 16, 32, 1),
    ("wgrad", 1, 8, 16, 64, 128, 1),
    ("wgrad", 1, 8, 16, 64, 64, 1),
    ("wgrad", 2, 19, 35, 128, 128, 1),
    ("wgrad", 2, 19, 35, 64, 128, 1),
    ("wgrad", 3, 20, 36, 32, 64, 2),
    ("wgrad", 2, 21, 37, 128, 128, 2),
    ("wgrad", 2, 19, 35, 16, 64, 1),
    ("wgrad", 1, 16, 32, 256, 256, 1),
    ("wgrad", 64, 97, 129, 128, 128, 1),
    # W-folded small-channel layers (W divisible by 4 / 2)
    ("fwd", 2, 19, 36, 16, 32, 1),
    ("fwd", 2, 19, 36, 32, 64, 1),
    ("fwd", 2, 13, 40, 8, 16, 1),
    ("dgrad", 2, 19, 36, 16, 32, 1),
    ("dgrad", 2, 19, 36, 32, 64, 1),
    ("dgrad", 2, 13, 40, 8, 16, 1),
    ("wgrad", 2, 19, 36, 16, 32, 1),
    ("wgrad", 2, 19, 36, 32, 64, 1),
    ("wgrad", 3, 13, 40, 8, 16, 1),
]
if os.environ.get("TC_ONLY"):
    CASES = [c for c in CASES if c[0] in os.environ["TC_ONLY"].split(",")]


def run_case(kind, N, H, W, Cin, Cout, s):
    import ctypes as C
    import torch
    from yogo_b200 import _lib as L

    lib = L.lib()
    dev = "cuda:0"
    g = torch.Generator().manual_seed(1)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5).to(dev)
    Ho, Wo = (H - 1) // s + 1, (W - 1) // s + 1
    res = {}
    if kind == "fwd":
        x = torch.randn(N, H, W, Cin, generator=g).to(dev).bfloat16()
        outs = {}
        for impl in ("simt", "tcgen05"):
            L.set_conv_impl(impl)
            y = torch.full((N, Ho, Wo, Cout), float("nan"), device=dev, dtype=torch.bfloat16)
            ep = L.FwdEpilogue(None, None, 0, None, None, None)
            L.check(lib.yg_conv_fwd(x.data_ptr(), w.data_ptr(), y.data_ptr(), 1, N, H, W, Cin, Cout, 3, s, C.byref(ep), L.stream()))
            torch.cuda.synchronize()
            outs[impl] = y.float()
    elif kind == "wgrad":
        x = torch.randn(N, H, W, Cin, generator=g).to(dev).bfloat16()
        dz = (torch.randn(N, Ho, Wo, Cout, generator=g) / (N * Ho * Wo) ** 0.5).to(dev).bfloat16()
        outs = {}
        import time
        for impl in ("simt", "tcgen05"):
            L.set_conv_impl(impl)
            dw = torch.full((Cout, Cin, 3, 3), float("nan"), device=dev)
            db = torch.full((Cout,), float("nan"), device=dev)
            nb = lib.yg_conv_wgrad_workspace(N, H, W, Cin, Cout, 3, s)
            ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
            for rep in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                L.check(lib.yg_conv_wgrad(x.data_ptr(), dz.data_ptr(), dw.data_ptr(), db.data_ptr(), 1, N, H, W, Cin, Cout, 3, s,
                                          0.0, ws.data_ptr(), nb, L.stream()))
                torch.cuda.synchronize()
                res["ms_" + impl] = round((time.perf_counter() - t0) * 1e3, 3)
            outs[impl] = dw.permute(2, 3, 0, 1).contiguous()[None].reshape(1, 3, 3, Cout * Cin)
            outs[impl + "_db"] = db
        res["db_rel"] = float((outs["simt_db"] - outs["tcgen05_db"]).norm() / outs["simt_db"].norm())
    else:
        dz = torch.randn(N, Ho, Wo, Cout, generator=g).to(dev).bfloat16()
        outs = {}
        for impl in ("simt", "tcgen05"):
            L.set_conv_impl(impl)
            y = torch.full((N, H, W, Cin), float("nan"), device=dev, dtype=torch.bfloat16)
            L.check(lib.yg_conv_dgrad(dz.data_ptr(), w.data_ptr(), y.data_ptr(), 1, N, H, W, Cin, Cout, 3, s, None, L.stream()))
            torch.cuda.synchronize()
            outs[impl] = y.float()
    a, b = outs["simt"], outs["tcgen05"]
    nan = int(torch.isnan(b).sum())
    d = (a - torch.nan_to_num(b)).abs()
    rel = float(d.norm() / a.norm())
    res.update(rel=rel, nan=nan, maxabs=float(d.max()))
    if rel > 2e-2 or nan:
        bad = d > 0.05 * a.abs().max()
        res["bad_frac"] = float(bad.float().mean())
        res["bad_by_row"] = [round(float(v), 2) for v in bad.float().mean(dim=(0, 2, 3))[:24]]
        res["bad_by_col"] = [round(float(v), 2) for v in bad.float().mean(dim=(0, 1, 3))[:40]]
        res["bad_by_ch"] = [round(float(v), 2) for v in bad.float().mean(dim=(0, 1, 2))[:64]]
        res["sample_simt"] = [round(float(v), 3) for v in a[0, 0, 0, :8]]
        res["sample_tc"] = [round(float(v), 3) for v in b[0, 0, 0, :8]]
        # ratio test: is tc a scaled / partial sum of simt?
        res["dot_ratio"] = float((a * torch.nan_to_num(b)).sum() / (a * a).sum())
    print("RESULT " + json.dumps(res), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        args = sys.argv[1:]
        run_case(args[0], *[int(v) for v in args[1:]])
    else:
        for c in CASES:
            cmd = [sys.executable, os.path.abspath(__file__)] + [str(v) for v in c]
            try:
                p = subprocess.run(cmd, capture_output=True, text=True, timeout=180)
                out = [l for l in p.stdout.splitlines() if l.startswith("RESULT")]
                tail = (p.stderr or "").strip().splitlines()[-3:]
                print(c, "rc", p.returncode, out[0] if out else "NO RESULT", "" if out else tail, flush=True)
            except subprocess.TimeoutExpired:
                print(c, "TIMEOUT", flush=True)
