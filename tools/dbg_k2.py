import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
import test_gpu_parity as T
from yogo_b200 import _lib as L
L.set_conv_impl("auto")
for i in [int(v) for v in sys.argv[1].split(",")]:
    sh = T.SHAPES[i]
    try:
        T._conv_case(*sh, dtype=torch.float32, act=(sh[3] % 3), with_stats=True, seed=sum(sh))
        print(i, sh, "ok")
    except AssertionError as e:
        print(i, sh, "FAIL", str(e)[:80])
