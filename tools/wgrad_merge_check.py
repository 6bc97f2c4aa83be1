"""tcgen05 wgrad variants against the plain path: merged row taps (one N = 192 MMA; option bit 19 disables) and the dz tile
shared by both parity groups of a stride-2 tile (option bit 20 disables)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from yogo_b200 import _lib as L
lib = L.lib(); dev = "cuda:0"
BASE = 25 + 8192 + 16384
for (N, H, W, Cin, Cout, s) in [(2, 9, 11, 64, 128, 1), (3, 21, 37, 64, 128, 1), (64, 193, 258, 64, 128, 1), (3, 26, 35, 32, 64, 2),
                               (2, 20, 28, 128, 128, 2), (2, 21, 37, 64, 128, 2), (64, 386, 516, 32, 64, 2), (64, 193, 258, 128, 128, 2)]:
    g = torch.Generator().manual_seed(N + H)
    x = torch.randn(N, H, W, Cin, generator=g).to(dev).bfloat16()
    dz = torch.randn(N, (H - 1) // s + 1, (W - 1) // s + 1, Cout, generator=g).to(dev).bfloat16()
    res = {}
    for name, opt in (("rows", BASE + (1 << 19) + (1 << 20) + (1 << 17)), ("merged", BASE + (1 << 17))):
        lib.yg_set_tc_options(opt)
        dw = torch.zeros(Cout, Cin, 3, 3, device=dev); db = torch.zeros(Cout, device=dev)
        nb = lib.yg_conv_wgrad_workspace(N, H, W, Cin, Cout, 3, s); ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
        f = lambda: L.check(lib.yg_conv_wgrad(x.data_ptr(), dz.data_ptr(), dw.data_ptr(), db.data_ptr(), 1, N, H, W, Cin, Cout, 3, s, 0.0, ws.data_ptr(), nb, L.stream()))
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): f()
        e1.record(); torch.cuda.synchronize()
        res[name] = (dw.clone(), db.clone(), e0.elapsed_time(e1) / 5)
    lib.yg_set_tc_options(BASE)
    d = (res["rows"][0] - res["merged"][0]).abs().max().item()
    ref = res["rows"][0].abs().max().item()
    print({"shape": (N, H, W, Cin, Cout, s), "max_abs_diff": d, "max_abs": ref, "db_diff": (res["rows"][1] - res["merged"][1]).abs().max().item(),
           "ms_rows": round(res["rows"][2], 4), "ms_merged": round(res["merged"][2], 4)})
