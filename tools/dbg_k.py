import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from yogo_b200 import _lib as L
from oracle import yogo_oracle as O
lib = L.lib(); DEV = "cuda:0"
def rel(a, b): return float((a - b).norm() / b.norm())
for (N, H, W, Cin, Cout, s) in [(1, 10, 14, 128, 128, 1), (2, 11, 14, 96, 192, 2)]:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    keep = (torch.rand(N, Cout, generator=g) > 0.2).float() / 0.8
    xd = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    yref = F.conv2d(x.double(), w.double(), b.double(), stride=s, padding=1).float()
    Ho, Wo = yref.shape[-2:]
    wd, bd, kd = w.to(DEV), b.to(DEV), keep.to(DEV).contiguous()
    import itertools
    for act, usek, uses in itertools.product((0, 1, 2), (False, True), (False, True)):
        name = f"act{act} keep{int(usek)} stats{int(uses)}"
        y = torch.empty(N, Ho, Wo, Cout, device=DEV)
        stats = torch.zeros(2 * Cout, dtype=torch.float64, device=DEV)
        ep = L.FwdEpilogue(None, bd.data_ptr(), act, kd.data_ptr() if usek else None, stats.data_ptr() if uses else None, None)
        rc = lib.yg_conv_fwd(xd.data_ptr(), wd.data_ptr(), y.data_ptr(), 0, N, H, W, Cin, Cout, 3, s, C.byref(ep), L.stream())
        torch.cuda.synchronize()
        a = O._act(yref, {0: None, 1: "lrelu", 2: "silu"}[act])
        if usek: a = a * keep[:, :, None, None]
        print((Cin, Cout, s), name, "rc", rc, "err", rel(y.permute(0, 3, 1, 2).cpu(), a))
