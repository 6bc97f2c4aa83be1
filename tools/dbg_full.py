import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import yogo_b200
from yogo_b200 import _lib as L
import test_gpu_parity as T
fname, case = sys.argv[1], sys.argv[2]
z = np.load("/root/repo/tests/golden/" + fname)
for dt in (torch.float32, torch.bfloat16):
    net, o, loss, comps, errs = T._zoo_step(z, case, dt)
    print(dt, "out", T._rel(o, z[case + ".train.out"]), "ref bf16", float(z[case + ".bf16err.out"][0]), "loss", loss.item(), z[case + ".train.loss"][0])
    for k, v in errs.items():
        print("  %-20s %.5f  ref-bf16err %.5f  ref-fp32err %.2e gradnorm %.4g" % (k, v, float(z[case + ".bf16err.grad." + k][0]), float(z[case + ".fp32err.grad." + k][0]), float(z[case + ".train.gradnorm." + k][0])))
