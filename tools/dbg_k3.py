import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from yogo_b200 import _lib as L
from oracle import yogo_oracle as O
lib = L.lib(); DEV = "cuda:0"
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
N, H, W, Cin, Cout, s = 1, 10, 14, 128, 128, 1
for seed in (1, 285):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    keep = (torch.rand(N, Cout, generator=g) > 0.2).float() / 0.8
    xd = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    y64 = F.conv2d(x.double(), w.double(), b.double(), stride=s, padding=1)
    y32 = F.conv2d(x, w, b, stride=s, padding=1)
    print("seed", seed, "cpu fp32 conv vs double", rel(y32, y64))
    wd, bd, kd = w.to(DEV), b.to(DEV), keep.to(DEV).contiguous()
    for impl in ("simt", "auto"):
        L.set_conv_impl(impl)
        for act in (0, 2):
            y = torch.empty(N, H, W, Cout, device=DEV)
            stats = torch.zeros(2 * Cout, dtype=torch.float64, device=DEV)
            ep = L.FwdEpilogue(None, bd.data_ptr(), act, kd.data_ptr(), stats.data_ptr(), None)
            lib.yg_conv_fwd(xd.data_ptr(), wd.data_ptr(), y.data_ptr(), 0, N, H, W, Cin, Cout, 3, s, C.byref(ep), L.stream())
            torch.cuda.synchronize()
            a64 = O._act(y64, {0: None, 2: "silu"}[act]) * keep[:, :, None, None]
            a32 = O._act(y32, {0: None, 2: "silu"}[act]) * keep[:, :, None, None]
            print("   ", impl, "act", act, "vs double", rel(y.permute(0, 3, 1, 2).cpu(), a64), "vs cpu fp32", rel(y.permute(0, 3, 1, 2).cpu(), a32))
