#!/usr/bin/env python
"""Benchmark of the YOGO hot path (BASELINE.json metric: train img/s, 772x1032, fwd+bwd+loss).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores

A "step" is one data-parallel training step of `base_model` on a batch of 64 synthetic
772x1032 grayscale images per GPU: forward, YOGOLoss, backward, gradient all-reduce (N > 1) and
the fused AdamW update.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os

os.environ.setdefault("NCCL_DEBUG", "WARN")   # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "train_img_per_s"
UNIT = "img/s"
H, W, NUM_CLASSES = 772, 1032, 7
SY, SX = 97, 129

# conv FLOPs per image, SURVEY.md Appendix A / BASELINE.md 3
TRAIN_GF = {"base_model": 66.483, "silu_model": 66.483, "double_filters": 265.472}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md), polled through NVML
    every 5 ms from a thread (nvidia-smi -lms is too coarse for a sub-second timed region)."""

    def __init__(self, index: int):
        self.index = index
        self.sm, self.reasons = [], set()
        self.sm_max = None
        self._stop = threading.Event()
        self.thread = None
        self.err = None

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
            while not self._stop.is_set():
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                mask = int(get_reasons(h))
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(0.005)
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=2)
        out = {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.sm_max,
               "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        if self.err:
            out["error"] = self.err
        return out


# ------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (torch CPU convolutions + YOGOLoss) restated in oracle/
# ------------------------------------------------------------------------------------------
def cpu_reference_step_time(model_name: str, batch: int, steps: int, warmup: int):
    """Train step (forward + YOGOLoss + backward) of the oracle port on all host cores.
    Returns (img/s, seconds per step, cores)."""
    from oracle import yogo_oracle as O
    import yogo_b200

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    net = yogo_b200.YOGO((H, W), O.ANCHOR_W, O.ANCHOR_H, NUM_CLASSES, model_func=yogo_b200.get_model_func(model_name))
    sd = net.state_dict()
    blocks = O.blocks_from_state_dict(model_name, sd)
    leaves = []
    for b in blocks:
        for t in [b.weight, b.bias] + ([b.bn["weight"], b.bn["bias"]] if b.bn else []):
            if t is not None:
                t.requires_grad_(True)
                leaves.append(t)
    x = O.synth_images(batch).float()
    lab = O.synth_labels(batch)
    g = torch.Generator().manual_seed(5)
    keeps = [(torch.rand(batch, b.weight.shape[0], generator=g) >= b.p_drop).float() if b.p_drop > 0 else None
             for b in blocks]

    def step():
        for t in leaves:
            t.grad = None
        tt = O.backbone_forward(x, blocks, train=True, drop_keep=keeps)
        out = O.head_transform(tt, O.ANCHOR_W, O.ANCHOR_H)
        _, _, dpred = O.yogo_loss_np(out.detach().numpy(), lab.numpy())
        out.backward(torch.from_numpy(dpred))

    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return batch / sec, sec, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_batch = 8
    v, sec, cores = cpu_reference_step_time(args.model, sample_batch, max(1, args.steps), max(1, args.warmup))
    sample = f"{sample_batch} images of the same 772x1032 workload per step, fp32, torch CPU ({cores} threads)"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.model} train step (fwd+YOGOLoss+bwd), 772x1032x1, 7 classes", "batch_per_step": sample_batch},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import yogo_b200
    from yogo_b200 import _lib as L
    from yogo_b200.train import DataParallelTrainer
    from tools import synth as O  # seeded synthetic inputs (shared with the tests); the oracle is not imported on this arm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    torch.manual_seed(0)
    net = yogo_b200.YOGO((H, W), O.ANCHOR_W, O.ANCHOR_H, NUM_CLASSES,
                         model_func=yogo_b200.get_model_func(args.model)).to(dev)
    net.compute_dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    net.train()
    loss_fn = yogo_b200.YOGOLoss().to(dev)
    trainer = DataParallelTrainer(net, loss_fn, total_steps=10000)
    trainer.broadcast_state()
    use_graph = bool(args.graph)   # several ranks: graph = fwd + loss + bwd, then one NCCL all-reduce and the fused AdamW

    nbuf = 2  # distinct host batches, alternated
    host_imgs = [O.synth_images(B, seed=10 * rank + i).pin_memory() for i in range(nbuf)]
    host_labs = [O.synth_labels(B, seed=100 + 10 * rank + i).pin_memory() for i in range(nbuf)]
    dev_imgs = [t.to(dev) for t in host_imgs]
    dev_labs = [t.to(dev) for t in host_labs]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value)
    if use_graph:
        trainer.step(dev_imgs[0], dev_labs[0])
        trainer.enable_cuda_graph(dev_imgs[0], dev_labs[0])
    for i in range(args.warmup):
        trainer.step(dev_imgs[i % nbuf], dev_labs[i % nbuf])
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.load().yg_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = trainer.step(dev_imgs[i % nbuf], dev_labs[i % nbuf])
    e1.record()
    barrier()
    launches = L.load().yg_launch_count() - launches0
    if use_graph:
        launches = (trainer.graph_launches + (0 if world == 1 else 1)) * args.steps   # (+ AdamW outside the graph)
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)
    final_loss = float(loss.item())

    # ---- end-to-end through the public API with HOST buffers
    # (host batches are fed through yogo_b200.train.DevicePrefetcher: the copy of step i+1 overlaps the compute of step i,
    # every step's inputs still cross PCIe inside the timed region and every step's loss is read back)
    from yogo_b200.train import DevicePrefetcher
    out_host = torch.empty(4, dtype=torch.float32).pin_memory()

    def e2e_run(nsteps):
        feeder = DevicePrefetcher(((host_imgs[i % nbuf], host_labs[i % nbuf]) for i in range(nsteps)), dev)
        for x, y in feeder:
            l = trainer.step(x, y)
            out_host[0:1].copy_(l.detach().reshape(1), non_blocking=True)

    e2e_run(max(1, args.warmup // 2))
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(t.item())
    h2d = host_imgs[0].numel() * host_imgs[0].element_size() + host_labs[0].numel() * host_labs[0].element_size()

    # ---- per-kernel breakdown with CUDA events (instrumented pass, not part of `value`)
    roof = None
    # every rank runs the instrumented steps (they contain collectives); rank 0 reports them
    saved_graph, trainer._graph = trainer._graph, None  # the breakdown needs individual launches
    breakdown, calls_per_step = kernel_breakdown(trainer, dev_imgs, dev_labs, args, L)
    trainer._graph = saved_graph
    barrier()
    if rank == 0:
        roof = roofline_from_breakdown(breakdown, calls_per_step, args, B)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, sec, cores = cpu_reference_step_time(args.model, 8, 3, 1)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "8 images/step of the same workload, fp32 torch-CPU port of the reference path (oracle/), 3 timed steps"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {
                "workload": f"{args.model} training step: fwd + YOGOLoss + bwd + grad all-reduce + fused AdamW, "
                            f"772x1032x1 uint8 images, 7 classes, batch {B}/GPU "
                            + ("(BASELINE configs[1])" if args.model == "base_model" and B == 64 else
                               "(BASELINE configs[3])" if args.model in ("double_filters", "silu_model") and B == 128 else
                               "(parity / sweep configuration)"),
                "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                "l2_policy": "inputs+activations per step (>3 GB) far exceed the 126 MB L2; batches alternate",
                "conv_impl": L.get_conv_impl(), "cuda_graph": bool(use_graph),
            },
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": cpu,
            "final_loss": final_loss,
            "kernel_breakdown_ms": breakdown,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def kernel_breakdown(trainer, dev_imgs, dev_labs, args, L):
    """Times every C-ABI call of a few steps with CUDA events on the launching stream."""
    lib = L.lib()
    names = [n for n in L.EXPORTED_SYMBOLS if n.startswith(("yg_conv", "yg_bn_act", "yg_bn_bwd", "yg_bn_stats", "yg_head", "yg_yogo", "yg_adamw"))
             and not n.endswith("workspace")]
    records = []
    originals = {}

    def wrap(name, fn):
        def inner(*a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*a)
            e1.record()
            records.append((name, a, e0, e1))
            return rc
        return inner

    class Proxy:
        def __getattr__(self, k):
            f = getattr(lib, k)
            return wrap(k, f) if k in names else f

    proxy = Proxy()
    orig_lib = L.lib
    L.lib = lambda: proxy
    try:
        steps = 2
        for i in range(steps):
            trainer.step(dev_imgs[i % len(dev_imgs)], dev_labs[i % len(dev_labs)])
        torch.cuda.synchronize()
    finally:
        L.lib = orig_lib
    agg = {}
    for name, a, e0, e1 in records:
        key = name
        if name in ("yg_conv_fwd", "yg_conv_dgrad", "yg_conv_wgrad"):
            # (.., dtype, N, H, W, Cin, Cout, k, s, ..): identify the layer by shape
            ints = [v for v in a if isinstance(v, int) and not isinstance(v, bool)]
            shape = [v for v in ints if 0 < v < 100000][:9]
            key = f"{name}[{'x'.join(str(v) for v in shape[-7:])}]"
        d = agg.setdefault(key, [0.0, 0])
        d[0] += e0.elapsed_time(e1) / steps
        d[1] += 1
    ordered = sorted(agg.items(), key=lambda kv: -kv[1][0])
    return ({k: round(v[0], 4) for k, v in ordered}, {k: v[1] / steps for k, v in ordered})


def measured_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture (profiles/ncu_traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum), or None when that kernel has not been captured."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            ent = json.load(f).get(kernel)
        return float(ent["dram_bytes_per_launch"]) if ent else None
    except (OSError, ValueError, KeyError):
        return None


def roofline_from_breakdown(breakdown, calls_per_step, args, B):
    """roofline of the dominant kernel (largest share of the step).  Convolutions are judged against the roof that
    binds them: arithmetic intensity (FLOP per algorithmic byte) above the machine balance -> tensor, else HBM
    (base_model layers 2, 3 are HBM-bound, 4-7 tensor-bound)."""
    peaks = load_peaks()
    if not breakdown:
        return None
    top, ms = next(iter(breakdown.items()))
    ms = ms / max(calls_per_step.get(top, 1.0), 1.0)   # the breakdown sums same-shaped layers: per launch here
    traffic = measured_traffic(top)
    if top.startswith(("yg_conv_fwd[", "yg_conv_dgrad[", "yg_conv_wgrad[")):
        dims = [int(v) for v in top[top.index("[") + 1:-1].split("x")]
        N, Hh, Ww, Cin, Cout, k, s = dims
        pad = k // 2
        Ho, Wo = (Hh + 2 * pad - k) // s + 1, (Ww + 2 * pad - k) // s + 1
        flops = 2.0 * N * Ho * Wo * Cout * Cin * k * k
        # algorithmic bytes: every activation tensor the op must touch once (bf16); dgrad also reads what the fused
        # activation backward of the previous layer needs
        nbytes = 2.0 * (N * Hh * Ww * Cin + N * Ho * Wo * Cout)
        if top.startswith("yg_conv_dgrad["):
            nbytes += N * Hh * Ww * Cin / 8.0   # + the 1-bit-per-element activation sign mask of the previous layer
        balance = peaks["bf16_tflops_sustained"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
        if flops / nbytes >= balance:
            ach = flops / (ms * 1e-3) / 1e12
            return {"kernel": top, "bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops_sustained"],
                    "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops_sustained"], "traffic": traffic,
                    "peak_source": peaks["source"] + " (sustained bf16, kernel timed inside a long step)",
                    "ms_per_launch": ms, "algorithmic_flops": flops}
        ach = nbytes / (ms * 1e-3) / 1e9
        return {"kernel": top, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peaks["source"] + " (copy bandwidth)",
                "ms_per_launch": ms, "algorithmic_bytes": nbytes, "flop_per_byte": flops / nbytes}
    # the other kernels of the step stream their operands once: algorithmic bytes per launch at the bench geometry
    H, W = 772, 1032
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    px1, px7 = B * Ho * Wo, B * 97 * 129
    stream_bytes = {
        "yg_conv_first_fwd": B * H * W + px1 * 16 * 2,              # uint8 image in, 16-channel bf16 out
        "yg_conv_first_bwd": B * H * W + px1 * 16 * 2,              # image + d(activation) in, 16x9 sums out
        "yg_conv_first_gram": B * H * W,
        "yg_bn_act_apply": 2 * px7 * 128 * 2 + px7 * 16,            # y_raw in, activation + sign mask out
        "yg_bn_bwd_apply": 3 * px7 * 128 * 2,                       # g, y_raw in, dz out
        "yg_bn_bwd_sums": 2 * px7 * 128 * 2,
        "yg_bn_stats": px7 * 128 * 2,
    }
    if top in stream_bytes:
        ach = stream_bytes[top] / (ms * 1e-3) / 1e9
        return {"kernel": top, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peaks["source"] + " (copy bandwidth)",
                "ms_per_launch": ms, "algorithmic_bytes": float(stream_bytes[top])}
    return {"kernel": top, "bound": "hbm", "achieved": None, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": None,
            "traffic": traffic, "ms_per_launch": ms, "peak_source": peaks["source"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="base_model")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", type=int, default=1, help="replay the whole step from a CUDA graph (single GPU)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
