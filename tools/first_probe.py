"""One launch of each first-layer kernel at the bench geometry (for ncu)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yogo_b200 import _lib as L
from tools import synth as S
dev = "cuda:0"
lib = L.lib()
N, H, W = 64, 772, 1032
Ho, Wo = 386, 516
x = S.synth_images(N).to(dev)
w = (torch.randn(16, 1, 3, 3) / 300).to(dev)
sc = torch.rand(16, device=dev) + 0.5
sh = torch.randn(16, device=dev) * 0.1
y = torch.empty(N, Ho, Wo, 16, device=dev, dtype=torch.bfloat16)
gram = torch.zeros(54, dtype=torch.float64, device=dev)
da = torch.randn(N, Ho, Wo, 16, device=dev).bfloat16()
P = torch.empty(16 * 9, device=dev); Sg = torch.empty(16, device=dev)
nb = lib.yg_conv_first_bwd_workspace(1, 16)
ws = torch.empty(nb, dtype=torch.uint8, device=dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
def run():
    ep = L.FwdEpilogue(sc.data_ptr(), sh.data_ptr(), 1, None, None, None, None)
    L.check(lib.yg_conv_first_fwd(x.data_ptr(), L.YG_U8, w.data_ptr(), y.data_ptr(), 1, N, H, W, 1, 16, 2, C.byref(ep), L.stream()))
    L.check(lib.yg_conv_first_gram(x.data_ptr(), L.YG_U8, N, H, W, 2, gram.data_ptr(), L.stream()))
    be = L.BwdEpilogue(None, 1, None, sc.data_ptr(), sh.data_ptr(), None, None, None, None)
    L.check(lib.yg_conv_first_bwd(x.data_ptr(), L.YG_U8, w.data_ptr(), da.data_ptr(), 1, N, H, W, 1, 16, 2, C.byref(be), None, None, None,
                                  P.data_ptr(), Sg.data_ptr(), 0.0, ws.data_ptr(), nb, L.stream()))
run(); torch.cuda.synchronize()
if reps > 1:
    for name, i in (("fwd", 0), ("gram", 1), ("bwd", 2)):
        pass
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e[0].record()
    for _ in range(reps): run()
    e[1].record(); torch.cuda.synchronize()
    print("fwd+gram+bwd ms per rep", e[0].elapsed_time(e[1]) / reps)
