"""world_size-2 gloo tests (CPU) of the data-parallel host logic in yogo_b200/train.py: flat bucket
layout, gradient-ready ordering, all-reduce semantics mean_over_ranks(clamp(local_grad)) (SURVEY.md
2.4: the clamp hook runs before DDP's reducer), initial broadcast and BN-buffer sync."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import yogo_b200
        from yogo_b200.train import DataParallelTrainer, FIRST_BUCKET_BYTES

        torch.manual_seed(100 + rank)  # deliberately different initial weights per rank
        net = yogo_b200.YOGO((64, 96), 0.05, 0.05, 7)
        tr = DataParallelTrainer(net, overlap=False)
        assert tr.world == world
        # layout: head first, then blocks last->first; first bucket closes at >= 1 MiB
        order = tr._ready_order(net._get_runner())
        assert order[0] is net.model[-1].weight and order[-2] is net.model[0][1].weight
        assert tr.buckets[0][0] == 0 and tr.buckets[-1][1] == tr.numel == net.num_params()
        assert (tr.buckets[0][1] - tr.buckets[0][0]) * 4 >= FIRST_BUCKET_BYTES
        assert all(a[1] == b[0] for a, b in zip(tr.buckets, tr.buckets[1:]))
        for p in net.parameters():
            o, k = tr.offsets[id(p)]
            assert p.data_ptr() == tr.flat_p[o:o + k].data_ptr() and p.grad.data_ptr() == tr.flat_g[o:o + k].data_ptr()
        # C1: broadcast makes every rank equal to rank 0
        tr.broadcast_state()
        ref = tr.flat_p.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(ref, tr.flat_p)
        # C3: each rank writes clamp(local_grad) into its views in ready order; buckets fire
        clip = 1.0
        g = torch.Generator().manual_seed(7 + rank)
        local = {}
        tr._bucket_pending = [0] * len(tr.buckets)
        fired = []
        orig = tr._launch_bucket
        tr._launch_bucket = lambda bi: (fired.append(bi), orig(bi))
        for p in order:
            lg = 4.0 * torch.randn(p.shape, generator=g)
            local[id(p)] = lg
            p.grad.copy_(lg.clamp(-clip, clip))
            tr._on_grad_ready(p)
        assert fired == list(range(len(tr.buckets)))
        for p in order:
            gathered = [torch.empty_like(local[id(p)]) for _ in range(world)]
            dist.all_gather(gathered, local[id(p)])
            expect = sum(t.clamp(-clip, clip) for t in gathered) / world
            torch.testing.assert_close(p.grad / world, expect)  # 1/world is folded into AdamW's grad_scale
        # C2-equivalent on demand
        with torch.no_grad():
            net.model[0][1].running_mean.fill_(float(rank + 1))
        tr.sync_bn_buffers()
        assert float(net.model[0][1].running_mean[0]) == 1.0
        # cosine schedule end points (train.py:219-223)
        tr.total_steps, tr.lr0, tr.decay_factor = 100, 3e-4, 10
        assert abs(tr.lr_at(0) - 3e-4) < 1e-12 and abs(tr.lr_at(100) - 3e-5) < 1e-12
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_data_parallel_host_logic_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
