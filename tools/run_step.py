"""Runs a few eager training steps of base_model at the bench configuration (for ncu captures)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import yogo_b200  # noqa: E402
from yogo_b200.train import DataParallelTrainer  # noqa: E402
from tools import synth as O  # noqa: E402  (seeded synthetic inputs)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = "cuda:0"
torch.manual_seed(0)
net = yogo_b200.YOGO((772, 1032), O.ANCHOR_W, O.ANCHOR_H, 7).to(dev)
net.train()
tr = DataParallelTrainer(net, yogo_b200.YOGOLoss().to(dev), total_steps=100)
img = O.synth_images(B).to(dev)
lab = O.synth_labels(B).to(dev)
for _ in range(steps):
    loss = tr.step(img, lab)
torch.cuda.synchronize()
print("loss", float(loss))
