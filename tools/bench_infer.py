"""Inference sweep (BASELINE config 5): eval forward (inference=True, BN folded) + objectness threshold + NMS +
per-class counts, batch 1..256, with CUDA events.  Two prediction regimes:
  dense  = random-init network output (~all 12,513 cells are candidates: adversarial for NMS)
  sparse = synthetic prediction tensors with K objects/image (realistic)

    python tools/bench_infer.py [--batches 1,8,64,256] [--reps 5]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import yogo_b200  # noqa: E402
from oracle import yogo_oracle as O  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="1,8,64,256")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--model", default="base_model")
    args = ap.parse_args()
    dev = "cuda:0"
    torch.manual_seed(0)
    net = yogo_b200.YOGO((772, 1032), O.ANCHOR_W, O.ANCHOR_H, 7, inference=True,
                         model_func=yogo_b200.get_model_func(args.model)).to(dev)
    net.eval()
    for B in [int(b) for b in args.batches.split(",")]:
        img = O.synth_images(B).to(dev)
        with torch.no_grad():
            fwd_ms = timed(lambda: net(img), args.reps)
            pred = net(img)
        dense_ms = timed(lambda: yogo_b200.format_preds_batch(pred), max(1, args.reps // 2))
        _, kc, _, _ = yogo_b200.format_preds_batch(pred)
        sp = O.synth_sparse_preds(B, K=300).to(dev)
        sparse_ms = timed(lambda: yogo_b200.format_preds_batch(sp), args.reps)
        _, kcs, _, counts = yogo_b200.format_preds_batch(sp)
        rec = {"batch": B, "fwd_ms": round(fwd_ms, 4), "fwd_img_s": round(B / fwd_ms * 1e3, 1),
               "nms_sparse_ms": round(sparse_ms, 4), "kept_sparse_per_img": float(kcs.float().mean()),
               "infer_img_s_sparse": round(B / (fwd_ms + sparse_ms) * 1e3, 1),
               "nms_dense_ms": round(dense_ms, 4), "kept_dense_per_img": float(kc.float().mean()),
               "infer_img_s_dense": round(B / (fwd_ms + dense_ms) * 1e3, 1),
               "nms_read_GBps_sparse": round(B * 600624 / sparse_ms / 1e6, 1)}
        print(json.dumps(rec), flush=True)
    # CPU reference (oracle port of format_preds + torchvision-semantics NMS) on a bounded sample
    sp = O.synth_sparse_preds(8, K=300)
    t0 = time.perf_counter()
    O.prediction_class_counts_np(sp.numpy())
    print(json.dumps({"cpu_oracle_nms_sparse_img_s": round(8 / (time.perf_counter() - t0), 1)}), flush=True)


if __name__ == "__main__":
    main()
