"""Experiment: can a 3x3 column tap be expressed as a +-1 pixel (one swizzled row) offset of the UMMA descriptor start
address inside ONE halo box, instead of a separately loaded, column-shifted box?  Compares interior tile columns."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from yogo_b200 import _lib as L  # noqa: E402

lib = L.lib()
dev = "cuda:0"
for (N, H, W, Cin, Cout) in [(1, 16, 32, 64, 64), (2, 24, 48, 128, 128), (1, 16, 32, 32, 64)]:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, H, W, Cin, generator=g).to(dev).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5).to(dev)
    outs = {}
    for name, opt in (("ref", 25), ("shift", 25 + 4096)):
        lib.yg_set_tc_options(opt)
        y = torch.zeros(N, H, W, Cout, device=dev, dtype=torch.bfloat16)
        ep = L.FwdEpilogue(None, None, 0, None, None, None, None)
        L.check(lib.yg_conv_fwd(x.data_ptr(), w.data_ptr(), y.data_ptr(), 1, N, H, W, Cin, Cout, 3, 1, C.byref(ep), L.stream()))
        torch.cuda.synchronize()
        outs[name] = y.float()
    lib.yg_set_tc_options(25)
    a, b = outs["ref"], outs["shift"]
    cols = torch.arange(W, device=dev)
    interior = ((cols % 16) >= 1) & ((cols % 16) <= 14)
    d_int = (a[:, :, interior] - b[:, :, interior]).abs().max().item()
    d_all = (a - b).abs().max().item()
    print({"shape": (N, H, W, Cin, Cout), "max_abs_diff_interior": d_int, "max_abs_diff_all": d_all, "ref_absmax": a.abs().max().item()})
