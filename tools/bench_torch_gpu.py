"""The reference's own GPU path on this box, for comparison ("the Blackwell kernel set to beat"): the same
`model_defns` backbone as ordinary torch.nn modules (what /root/reference/yogo/model.py builds) run by stock PyTorch - cuDNN
convolutions under bf16 autocast, channels_last, cudnn.benchmark - forward + backward with a stand-in scalar loss (the backbone is
> 98 % of the reference's step), and the eval forward.  None of this repository's kernels run here.

    python tools/bench_torch_gpu.py [--model base_model] [--batch 64] [--steps 20]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import yogo_b200  # noqa: E402  (only for the torch.nn module tree and its default init)
from tools import synth as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="base_model")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    net = yogo_b200.YOGO((772, 1032), S.ANCHOR_W, S.ANCHOR_H, 7, model_func=yogo_b200.get_model_func(args.model))
    backbone = net.model.to(dev).to(memory_format=torch.channels_last)
    x = S.synth_images(args.batch).to(dev).float().contiguous(memory_format=torch.channels_last)
    params = [p for p in backbone.parameters()]

    def train_step():
        for p in params:
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = backbone(x)
        out.float().square().mean().backward()

    def infer_step():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            backbone(x)

    res = {"model": args.model, "batch": args.batch, "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}
    for name, fn, train in (("train_fwd_bwd", train_step, True), ("eval_fwd", infer_step, False)):
        backbone.train(train)
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        res[name + "_ms"] = round(ms, 3)
        res[name + "_img_s"] = round(args.batch / ms * 1e3, 1)
    # post-processing as the reference does it (utils/prediction_formatting.py:23-93, infer.py:60-87) with torch / torchvision
    # ops on the GPU: one Python iteration per image - mask gather, box_convert, ops.nms (CUDA), fancy index, argmax counts
    import torchvision.ops as ops
    preds = S.synth_sparse_preds(args.batch, K=300).to(dev)

    def post():
        tot = torch.zeros(7, dtype=torch.long, device=dev)
        for pred in preds:
            p = pred.view(pred.shape[0], -1).T
            c = p[p[:, 4] > 0.5]
            keep = ops.nms(ops.box_convert(c[:, :4], "cxcywh", "xyxy"), c[:, 5:].max(dim=1).values * c[:, 4], 0.5)
            rows = c[keep]
            vals, idx = rows[:, 5:].max(dim=1)
            tot += torch.nn.functional.one_hot(idx[vals > 0], num_classes=7).sum(dim=0)
        return tot

    for _ in range(3):
        post()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        post()
    e1.record()
    torch.cuda.synchronize()
    res["torch_postproc_sparse_ms"] = round(e0.elapsed_time(e1) / 5, 3)
    rows, kc, _, counts = yogo_b200.format_preds_batch(preds)
    assert torch.equal(counts.cpu(), post().cpu())   # same per-class counts
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
