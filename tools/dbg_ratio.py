import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import test_gpu_parity as T
from _zoo_cases import ZOO_CASES
for fname, cases in (("model_zoo.npz", ZOO_CASES), ("model_full.npz", T.FULL_CASES)):
    z = np.load("/root/repo/tests/golden/" + fname)
    for case in cases:
        net, o, loss, comps, errs = T._zoo_step(z, case, torch.bfloat16)
        r = {k: v / max(float(z[case + ".bf16err.grad." + k][0]), 1e-2) for k, v in errs.items()}
        print("%-20s out %.2f | " % (case, T._rel(o, z[case + ".train.out"]) / float(z[case + ".bf16err.out"][0])) + " ".join("%s:%.2f" % (k.replace("model.", ""), v) for k, v in r.items()))
