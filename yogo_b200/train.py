"""Data-parallel training step: the hot loop of /root/reference/yogo/train.py:295-339 and the
DDP/optimizer wiring of train.py:152-159, 206-223, re-designed for one process per B200.

What the reference does with ``DistributedDataParallel`` + ``torch.optim.AdamW``:
  * DDP ctor broadcast of parameters/buffers from rank 0 (C1)            -> ``broadcast_state``
  * bucketed gradient all-reduce (SUM, / world) overlapped with backward (C3)
                                                                          -> flat fp32 buckets,
    the weight-gradient kernels write (already clamped) gradients straight into bucket views,
    ``on_grad_ready`` fires as each layer's gradient is enqueued (last layers first), and a
    complete bucket is all-reduced on a side stream while the remaining dgrad/wgrad kernels run
  * per-step buffer broadcast (C2)                                        -> dropped: only BN
    running statistics change, and they are broadcast from rank 0 on demand
    (``sync_bn_buffers``) before eval/checkpoint, which is result-equivalent because rank 0's
    statistics are what the reference checkpoints (train.py:273-288)
  * AdamW + CosineAnnealingLR (train.py:213-223)                          -> one fused kernel over
    the flat parameter buffer (csrc/optim.cu), 1/world folded into the gradient scale.
Semantics preserved: mean over ranks of clamp(local_grad, +-clip), loss divided
by the local batch, BatchNorm with local statistics (no SyncBN).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib as L
from .model import YOGO
from .yogo_loss import YOGOLoss

FIRST_BUCKET_BYTES = 1 << 20   # DDP: first bucket 1 MiB ...
BUCKET_BYTES = 25 << 20        # ... then bucket_cap_mb = 25


class DataParallelTrainer:
    def __init__(
        self,
        net: YOGO,
        loss_fn: Optional[YOGOLoss] = None,
        lr: float = 3e-4,
        weight_decay: float = 5e-2,
        betas: Tuple[float, float] = (0.9, 0.999),
        eps: float = 1e-8,
        total_steps: Optional[int] = None,
        decay_factor: float = 10.0,
        process_group=None,
        overlap: bool = True,
    ):
        self.net = net
        self.loss_fn = loss_fn if loss_fn is not None else YOGOLoss()
        self.lr0, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.total_steps, self.decay_factor = total_steps, decay_factor
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.overlap = overlap
        self.step_count = 0
        self._graph = None

        runner = net._get_runner()
        params = [p for p in runner.plan.params if p.requires_grad]
        self.params = params
        dev = params[0].device
        n = sum(p.numel() for p in params)
        self.numel = n
        # flat parameter / gradient / optimizer-state buffers; parameters become views
        self.flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        # bucket layout follows gradient-ready order: head first, then blocks last -> first
        order = self._ready_order(runner)
        off = 0
        self.offsets: Dict[int, Tuple[int, int]] = {}
        for p in order:
            self.offsets[id(p)] = (off, p.numel())
            off += p.numel()
        assert off == n
        views = {}
        with torch.no_grad():
            for p in order:
                o, k = self.offsets[id(p)]
                self.flat_p[o:o + k].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[o:o + k].view_as(p)
                g = self.flat_g[o:o + k].view_as(p)
                p.grad = g
                views[id(p)] = g
        runner.grad_views = views
        # buckets: contiguous ranges of the flat buffer
        self.buckets: List[Tuple[int, int]] = []
        self.bucket_of: Dict[int, int] = {}
        start, cap = 0, FIRST_BUCKET_BYTES
        cur = 0
        for p in order:
            o, k = self.offsets[id(p)]
            self.bucket_of[id(p)] = len(self.buckets)
            cur = o + k
            if (cur - start) * 4 >= cap:
                self.buckets.append((start, cur))
                start, cap = cur, BUCKET_BYTES
        if cur > start:
            self.buckets.append((start, cur))
        self._bucket_pending = [0] * len(self.buckets)
        self._bucket_size = [0] * len(self.buckets)
        for p in order:
            self._bucket_size[self.bucket_of[id(p)]] += 1
        self.comm_stream = torch.cuda.Stream(device=dev) if (self.world > 1 and dev.type == "cuda") else None
        self._comm_events: List[torch.cuda.Event] = []
        runner.on_grad_ready = self._on_grad_ready if self.world > 1 else None

    @staticmethod
    def _ready_order(runner) -> List[torch.nn.Parameter]:
        plan = runner.plan
        order: List[torch.nn.Parameter] = [plan.head.weight, plan.head.bias]
        for blk in reversed(plan.blocks):
            order.append(blk.conv.weight)
            if blk.conv.bias is not None:
                order.append(blk.conv.bias)
            if blk.bn is not None:
                order += [blk.bn.weight, blk.bn.bias]
        return [p for p in order if p.requires_grad]

    # ------------------------------------------------------------------ distributed plumbing
    def broadcast_state(self, src: int = 0) -> None:
        """C1: every rank starts from rank `src`'s parameters and buffers."""
        if self.world == 1:
            return
        dist.broadcast(self.flat_p, src, group=self.pg)
        for b in self.net.buffers():
            if b.is_floating_point() or b.dtype in (torch.int64, torch.bool):
                t = b if b.dtype != torch.bool else b.to(torch.uint8)
                dist.broadcast(t, src, group=self.pg)
                if b.dtype == torch.bool:
                    b.copy_(t.bool())

    def sync_bn_buffers(self, src: int = 0) -> None:
        """C2-equivalent, on demand: BN running statistics of rank `src` to every rank."""
        if self.world == 1:
            return
        for name, b in self.net.named_buffers():
            if "running_" in name or "num_batches_tracked" in name:
                dist.broadcast(b, src, group=self.pg)

    def _on_grad_ready(self, p) -> None:
        if getattr(self, "_defer_allreduce", False):
            return   # CUDA-graph mode with several ranks: one all-reduce of the whole flat buffer after the replay
        bi = self.bucket_of.get(id(p))
        if bi is None:
            return
        self._bucket_pending[bi] += 1
        if self._bucket_pending[bi] == self._bucket_size[bi]:
            self._launch_bucket(bi)

    def _launch_bucket(self, bi: int) -> None:
        s, e = self.buckets[bi]
        cur = torch.cuda.current_stream() if self.comm_stream is not None else None
        if self.overlap and self.comm_stream is not None:
            ev = torch.cuda.Event()
            ev.record(cur)
            self.comm_stream.wait_event(ev)
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(self.flat_g[s:e], op=dist.ReduceOp.SUM, group=self.pg)
                done = torch.cuda.Event()
                done.record(self.comm_stream)
            self._comm_events.append(done)
        else:
            dist.all_reduce(self.flat_g[s:e], op=dist.ReduceOp.SUM, group=self.pg)

    # ------------------------------------------------------------------ one training step
    def lr_at(self, step: int) -> float:
        """CosineAnnealingLR(T_max=total_steps, eta_min=lr/decay_factor) (train.py:219-223)."""
        if not self.total_steps:
            return self.lr0
        eta_min = self.lr0 / self.decay_factor
        t = min(step, self.total_steps)
        return eta_min + (self.lr0 - eta_min) * (1 + math.cos(math.pi * t / self.total_steps)) / 2

    def forward_backward(self, imgs: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        self._bucket_pending = [0] * len(self.buckets)
        self._comm_events = []
        out = self.net(imgs)
        loss, _ = self.loss_fn(out, labels)
        loss.backward()
        if self.world > 1:
            cur = torch.cuda.current_stream()
            for ev in self._comm_events:
                cur.wait_event(ev)
        return loss

    def optimizer_step(self) -> None:
        self.step_count += 1
        b1, b2 = self.betas
        L.check(L.lib().yg_adamw_flat(self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.exp_avg.data_ptr(),
                                      self.exp_avg_sq.data_ptr(), self.numel, self.lr_at(self.step_count - 1),
                                      b1, b2, self.eps, self.wd, self.step_count, 1.0 / self.world, L.stream()))

    def step(self, imgs: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """zero_grad is implicit: every gradient view is overwritten by its kernel each step."""
        if self._graph is not None:
            return self._graph_step(imgs, labels)
        loss = self.forward_backward(imgs, labels)
        self.optimizer_step()
        return loss

    # ------------------------------------------------------------------ CUDA-graph replay of the whole step
    def _hyper_host(self, step_count: int) -> torch.Tensor:
        b1, b2 = self.betas
        return torch.tensor([self.lr_at(step_count - 1), b1, b2, self.eps, self.wd, 1.0 - b1 ** step_count,
                             math.sqrt(1.0 - b2 ** step_count), 1.0 / self.world], dtype=torch.float32)

    def enable_cuda_graph(self, imgs: torch.Tensor, labels: torch.Tensor, warmup: int = 2) -> None:
        """Capture forward + loss + backward (+ all-reduce) + AdamW for this input shape into one CUDA graph;
        afterwards `step` copies the batch into static buffers and replays it (launch-bound inner loop:
        ~350 kernel launches and their host-side argument marshalling collapse into one graph launch)."""
        dev = imgs.device
        self._static_imgs = imgs.clone()
        self._static_labels = labels.clone()
        self._hyper = torch.zeros(8, dtype=torch.float32, device=dev)

        # several ranks: the graph holds forward + loss + backward only; the gradient all-reduce (one NCCL call on the whole
        # 2 MB flat buffer, a few tens of microseconds over NVLink) and the fused AdamW follow the replay on the same stream,
        # so no collective is ever captured
        self._graph_has_optimizer = self.world == 1
        self._defer_allreduce = self.world > 1

        def body():
            loss = self.forward_backward(self._static_imgs, self._static_labels)
            if self._graph_has_optimizer:
                self._adamw_dev()
            return loss

        # warm-up on a side stream (allocator pools, library caches), with a learning rate of zero so that the
        # parameters and optimizer state are not disturbed ... they are: so snapshot and restore instead
        snap = (self.flat_p.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone(),
                [b.clone() for b in self.net.buffers()])
        self._hyper.copy_(self._hyper_host(1))
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        n0 = L.load().yg_launch_count()
        with torch.cuda.graph(graph):
            self._graph_loss = body()
        self.graph_launches = int(L.load().yg_launch_count() - n0)  # our kernels inside one replay
        self.flat_p.copy_(snap[0]); self.exp_avg.copy_(snap[1]); self.exp_avg_sq.copy_(snap[2])
        for b, v in zip(self.net.buffers(), snap[3]):
            b.copy_(v)
        torch.cuda.synchronize()
        self._graph = graph

    def _adamw_dev(self) -> None:
        L.check(L.lib().yg_adamw_flat_dev(self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.exp_avg.data_ptr(),
                                          self.exp_avg_sq.data_ptr(), self.numel, self._hyper.data_ptr(), L.stream()))

    def _graph_step(self, imgs: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        self.step_count += 1
        self._hyper.copy_(self._hyper_host(self.step_count))  # 32 bytes from pageable memory: staged, race-free
        self._static_imgs.copy_(imgs, non_blocking=True)
        self._static_labels.copy_(labels, non_blocking=True)
        self._graph.replay()
        if not self._graph_has_optimizer:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.pg)
            self._adamw_dev()
        return self._graph_loss


class DevicePrefetcher:
    """Feeds pinned host batches to the GPU one step ahead of the consumer (what `DataLoader(pin_memory=True)` +
    `.to(device, non_blocking=True)` does in the reference's training loop, train.py:288-291, made explicit): the
    host->device copy of batch i+1 runs on a side stream while step i computes, so the copy leaves the critical path.
    Iterating yields `(images, labels)` device tensors that are safe to use on the current stream."""

    def __init__(self, batches, device):
        self.batches = iter(batches)
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self._next = None
        self._fetch()

    def _fetch(self):
        try:
            img, lab = next(self.batches)
        except StopIteration:
            self._next = None
            return
        with torch.cuda.stream(self.stream):
            x = img.to(self.device, non_blocking=True)
            y = lab.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._next = (x, y, ev)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        x, y, ev = self._next
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        x.record_stream(cur)
        y.record_stream(cur)
        self._fetch()   # the next copy overlaps with the step the caller is about to run
        return x, y
