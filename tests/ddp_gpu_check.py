"""Multi-rank GPU equivalence check of yogo_b200.train.DataParallelTrainer (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/ddp_gpu_check.py

What the reference gets from DistributedDataParallel (/root/reference/yogo/train.py:159, 322): after a step every rank
holds the same parameters, and they are the parameters a single process would get from the gradient
mean_r clamp(g_r, +-clip) (the clamp hook of model.py:76-77 runs before DDP's reducer).  Checked here for
 (1) the eager path (bucketed NCCL all-reduce on a side stream overlapped with backward),
 (2) the CUDA-graph path (the same bucketed all-reduces captured inside the graph, AdamW in the graph),
both against (3) a single-rank trainer on the same GPU fed every rank's shard in turn.
tests/test_gpu_parity.py::test_multi_rank_equivalence launches this file when the box has >= 2 GPUs.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import yogo_b200  # noqa: E402
from tools.synth import ANCHOR_H, ANCHOR_W, synth_fill_model, synth_images, synth_labels  # noqa: E402
from yogo_b200.train import DataParallelTrainer  # noqa: E402


def make(dev, pg, total_steps=20):
    net = yogo_b200.YOGO((196, 260), ANCHOR_W, ANCHOR_H, 7)
    synth_fill_model(net, 5)
    net = net.to(dev)
    net.compute_dtype = torch.bfloat16
    net.train()
    for m in net.model.modules():
        if isinstance(m, torch.nn.Dropout2d):
            m.p = 0.0
    return net, DataParallelTrainer(net, yogo_b200.YOGOLoss().to(dev), total_steps=total_steps, process_group=pg)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    solo = [dist.new_group([r]) for r in range(world)][rank]   # a world-1 group per rank for the single-process reference
    B = 4
    imgs = synth_images(B * world, 196, 260, seed=3).to(dev)
    Sy, Sx = yogo_b200.YOGO((196, 260), ANCHOR_W, ANCHOR_H, 7).get_grid_size()[::-1]
    labels = synth_labels(B * world, Sy, Sx, 7, 30, seed=4).to(dev)
    shard = slice(rank * B, (rank + 1) * B)

    def gathered_equal(t):
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t.contiguous())
        return all(torch.equal(out[0], o) for o in out)

    # ---- (3) single-process reference: gradient = mean over shards of the (already clamped) per-shard gradient
    net_r, tr_r = make(dev, solo)
    assert tr_r.world == 1
    bufs = [b.clone() for b in net_r.buffers()]
    gsum = torch.zeros_like(tr_r.flat_g)
    for r in range(world):
        tr_r.forward_backward(imgs[r * B:(r + 1) * B], labels[r * B:(r + 1) * B])
        assert float(tr_r.flat_g.abs().max()) <= 1.0   # clamp(+-clip_value) happened inside the gradient kernels
        gsum += tr_r.flat_g
        for b, v in zip(net_r.buffers(), bufs):
            b.copy_(v)
    tr_r.flat_g.copy_(gsum / world)
    g_ref = tr_r.flat_g.clone()
    tr_r.optimizer_step()

    # ---- (1) eager data-parallel step
    net_e, tr_e = make(dev, None)
    assert tr_e.world == world and len(tr_e.buckets) >= 1
    tr_e.broadcast_state()
    loss_e = tr_e.step(imgs[shard], labels[shard])
    torch.cuda.synchronize()
    assert gathered_equal(tr_e.flat_g), "all-reduced gradients differ between ranks"
    assert gathered_equal(tr_e.flat_p), "parameters differ between ranks after the eager step"
    rel_g = float((tr_e.flat_g / world - g_ref).norm() / g_ref.norm())
    assert rel_g < 1e-5, rel_g
    dp = (tr_e.flat_p - tr_r.flat_p).abs()
    frac_e = float((dp > 1e-7).float().mean())
    assert frac_e < 1e-3, frac_e     # Adam's first step is +-lr * sign(g): only g ~ 1e-8 can differ

    # ---- (2) CUDA-graph data-parallel steps: collectives captured, overlapped, AdamW inside
    net_g, tr_g = make(dev, None)
    tr_g.broadcast_state()
    tr_g.enable_cuda_graph(imgs[shard], labels[shard])
    assert tr_g.graph_launches > 20
    loss_g = tr_g.step(imgs[shard], labels[shard])
    torch.cuda.synchronize()
    assert gathered_equal(tr_g.flat_p), "parameters differ between ranks after the graph step"
    assert abs(loss_g.item() - loss_e.item()) <= 1e-4 * abs(loss_e.item()), (loss_g.item(), loss_e.item())
    frac_g = float(((tr_g.flat_p - tr_r.flat_p).abs() > 1e-7).float().mean())
    assert frac_g < 1e-3, frac_g
    for _ in range(3):
        tr_e.step(imgs[shard], labels[shard])
        lg = tr_g.step(imgs[shard], labels[shard])
    torch.cuda.synchronize()
    assert gathered_equal(tr_g.flat_p) and gathered_equal(tr_e.flat_p)
    assert torch.isfinite(lg).all()
    # after several steps Adam has turned sign flips of ~0 gradients (atomics order in the BN statistics) into +-lr
    # differences that propagate: compared statistically, as in test_trainer_step_matches_torch_adamw_and_graph_replay
    drift = float((tr_g.flat_p - tr_e.flat_p).norm() / tr_e.flat_p.norm())
    bad = float(((tr_g.flat_p - tr_e.flat_p).abs() > 1e-6 + 1e-5 * tr_e.flat_p.abs()).float().mean())
    assert drift < 5e-3, (drift, bad)   # (in bf16 most parameters differ in their last bits by then: `bad` is informational)
    # BN running statistics are per rank until asked for (C2 on demand)
    tr_g.sync_bn_buffers()
    assert gathered_equal(net_g.model[0][1].running_mean)
    if rank == 0:
        print("ddp_gpu_check ok: world %d, buckets %d, graph launches %d, grad rel err %.2e, param mismatch eager %.1e graph %.1e, "
              "4-step drift graph vs eager %.2e" % (world, len(tr_e.buckets), tr_g.graph_launches, rel_g, frac_e, frac_g, drift))
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)   # the step graphs hold captured NCCL work: skip the teardown of graphs / process group (it can block)


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        import traceback
        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)   # a failed rank must not wait in NCCL teardown for its peers
