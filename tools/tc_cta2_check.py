"""CTA-pair (cta_group::2) engine vs the single-CTA engine on the same inputs: forward and dgrad, with epilogues."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from yogo_b200 import _lib as L  # noqa: E402

lib = L.lib()
dev = "cuda:0"
OPT1, OPT2 = 25 + 8192, 25 + 8192 + 16384
for (N, H, W, Cin, Cout) in [(1, 16, 16, 128, 128), (2, 19, 35, 128, 128), (3, 33, 20, 128, 128), (64, 97, 129, 128, 128)]:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, H, W, Cin, generator=g).to(dev).bfloat16()
    dz = torch.randn(N, H, W, Cout, generator=g).to(dev).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5).to(dev)
    b = (torch.randn(Cout, generator=g) * 0.1).to(dev)
    res = {}
    for name, opt in (("one", OPT1), ("pair", OPT2)):
        lib.yg_set_tc_options(opt)
        y = torch.zeros(N, H, W, Cout, device=dev, dtype=torch.bfloat16)
        stats = torch.zeros(2 * Cout, dtype=torch.float64, device=dev)
        ep = L.FwdEpilogue(None, b.data_ptr(), 1, None, stats.data_ptr(), None, None)
        L.check(lib.yg_conv_fwd(x.data_ptr(), w.data_ptr(), y.data_ptr(), 1, N, H, W, Cin, Cout, 3, 1, C.byref(ep), L.stream()))
        dx = torch.zeros(N, H, W, Cin, device=dev, dtype=torch.bfloat16)
        be = L.BwdEpilogue(x.data_ptr(), 1, None, None, None, None, None, None, None)
        L.check(lib.yg_conv_dgrad(dz.data_ptr(), w.data_ptr(), dx.data_ptr(), 1, N, H, W, Cin, Cout, 3, 1, C.byref(be), L.stream()))
        torch.cuda.synchronize()
        t = []
        for fn in ("fwd", "dgrad"):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                if fn == "fwd":
                    L.check(lib.yg_conv_fwd(x.data_ptr(), w.data_ptr(), y.data_ptr(), 1, N, H, W, Cin, Cout, 3, 1, C.byref(ep), L.stream()))
                else:
                    L.check(lib.yg_conv_dgrad(dz.data_ptr(), w.data_ptr(), dx.data_ptr(), 1, N, H, W, Cin, Cout, 3, 1, C.byref(be), L.stream()))
            e1.record()
            torch.cuda.synchronize()
            t.append(round(e0.elapsed_time(e1) / 3, 4))
        res[name] = (y.float(), dx.float(), stats.clone(), t)
    lib.yg_set_tc_options(OPT1)
    a, p2 = res["one"], res["pair"]
    print({"shape": (N, H, W, Cin, Cout), "fwd_maxdiff": (a[0] - p2[0]).abs().max().item(), "dgrad_maxdiff": (a[1] - p2[1]).abs().max().item(),
           "ms_one": a[3], "ms_pair": p2[3]}, flush=True)

# stride-2 forward as CTA pairs (four parity boxes per K chunk)
for (N, H, W, Cin, Cout) in [(2, 21, 37, 128, 128), (64, 193, 258, 128, 128)]:
    g = torch.Generator().manual_seed(2)
    x = torch.randn(N, H, W, Cin, generator=g).to(dev).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5).to(dev)
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    outs = {}
    for name, opt in (("one", OPT1), ("pair", OPT2)):
        lib.yg_set_tc_options(opt)
        y = torch.zeros(N, Ho, Wo, Cout, device=dev, dtype=torch.bfloat16)
        ep = L.FwdEpilogue(None, None, 1, None, None, None, None)
        L.check(lib.yg_conv_fwd(x.data_ptr(), w.data_ptr(), y.data_ptr(), 1, N, H, W, Cin, Cout, 3, 2, C.byref(ep), L.stream()))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            L.check(lib.yg_conv_fwd(x.data_ptr(), w.data_ptr(), y.data_ptr(), 1, N, H, W, Cin, Cout, 3, 2, C.byref(ep), L.stream()))
        e1.record()
        torch.cuda.synchronize()
        outs[name] = (y.float(), round(e0.elapsed_time(e1) / 3, 4))
    lib.yg_set_tc_options(OPT2)
    print({"shape_s2": (N, H, W, Cin, Cout), "fwd_maxdiff": (outs["one"][0] - outs["pair"][0]).abs().max().item(),
           "ms_one": outs["one"][1], "ms_pair": outs["pair"][1]}, flush=True)

# stride-2 dgrad as CTA pairs (four parity classes, class-per-round tile order)
for (N, H, W, Cin, Cout) in [(2, 21, 37, 128, 128), (3, 20, 28, 128, 128), (64, 193, 258, 128, 128)]:
    g = torch.Generator().manual_seed(3)
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    dz = torch.randn(N, Ho, Wo, Cout, generator=g).to(dev).bfloat16()
    xs = torch.randn(N, H, W, Cin, generator=g).to(dev).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5).to(dev)
    outs = {}
    for name, opt in (("one", OPT1), ("pair", OPT2)):
        lib.yg_set_tc_options(opt)
        dx = torch.full((N, H, W, Cin), float("nan"), device=dev, dtype=torch.bfloat16)
        be = L.BwdEpilogue(xs.data_ptr(), 1, None, None, None, None, None, None, None)
        L.check(lib.yg_conv_dgrad(dz.data_ptr(), w.data_ptr(), dx.data_ptr(), 1, N, H, W, Cin, Cout, 3, 2, C.byref(be), L.stream()))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            L.check(lib.yg_conv_dgrad(dz.data_ptr(), w.data_ptr(), dx.data_ptr(), 1, N, H, W, Cin, Cout, 3, 2, C.byref(be), L.stream()))
        e1.record()
        torch.cuda.synchronize()
        outs[name] = (dx.float(), round(e0.elapsed_time(e1) / 3, 4))
    lib.yg_set_tc_options(OPT2)
    d = (outs["one"][0] - outs["pair"][0]).abs()
    print({"dgrad_s2": (N, H, W, Cin, Cout), "maxdiff": d.max().item(), "nan": int(torch.isnan(outs["pair"][0]).sum()),
           "ms_one": outs["one"][1], "ms_pair": outs["pair"][1]}, flush=True)
