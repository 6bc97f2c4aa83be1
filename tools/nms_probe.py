"""One format_preds_batch call per configuration (for an ncu launch list)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, yogo_b200
from tools import synth as S
dev = "cuda:0"
for B, kind in ((64, "sparse"), (256, "sparse"), (1, "dense"), (64, "dense")):
    if kind == "sparse":
        p = S.synth_sparse_preds(B, K=300).to(dev)
    else:
        g = torch.Generator().manual_seed(31)
        p = torch.rand(B, 12, 97, 129, generator=g)
        p[:, 2:4] = 0.02 + 0.05 * p[:, 2:4]
        p[:, 4] = 0.5 + 0.5 * p[:, 4] - 0.02
        p = p.to(dev)
    for _ in range(2):
        out = yogo_b200.format_preds_batch(p)
    torch.cuda.synchronize()
    print(B, kind, float(out[1].float().mean()))
