#!/usr/bin/env python
"""Benchmark of the volumetric-aggregation hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl reference]

A step = one pass of the hot path (feature packing + fused unproject+aggregate)
over one batch of synthetic input.  At N=1 the workload is BASELINE.json's
configs[1] (cfg2: B8 V4 C32 96x96 -> 64^3, softmax fusion, fp32).  For N>1 the
driver launches this file under torchrun, one rank per GPU; every rank owns its
own batch of the same shape (batch sharding, no collective in the data path,
"weak" scaling) and the reported value is the units of all ranks over the
slowest rank's device time.

Rank 0 prints ONE JSON line.  `value` is timed with inputs resident in HBM;
`e2e` goes through the public Python API (`unprojection` -> ctypes -> C ABI)
with pinned HOST buffers, host->device copies of every input and a
device->host read of the step's metric inside the timed region (median of
three repetitions).  `extras.channels_last_in_place` is the same device-resident
step when the feature maps arrive as (B,V,H,W,C) and are gathered in place
(informative, not the headline).

`--impl reference` times the reference's own CPU path (oracle/torch_port.py:
the same ATen calls in the same order, all host threads) on a bounded sample of
the same workload.  That is the only place besides `cpu_baseline` where this
file executes anything under oracle/.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from multiviewhmr_b200 import synthetic as syn  # noqa: E402

METRIC = "Gvoxel-ch-views/s fused unproject+aggregate"
UNIT = "Gvcv/s"
L2_BYTES = 126 * 1024 * 1024


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(workload):
    """dram read+write bytes per launch of the dominant kernel from the committed ncu capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh).get(workload)
    except Exception:
        return None


class ClockSampler:
    """SM clock + throttle reasons sampled every few ms while the timed region runs."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, dev):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:        # NVML numbering ignores CUDA_VISIBLE_DEVICES: resolve by UUID
                uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(dev.index or 0)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv is not None:
            self._stop.clear()
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
            self._thread = None

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def cpu_reference_sample(w, steps, warmup, n_samples=2):
    """Time the torch port of the reference on the host cores: the first
    `n_samples` samples of the workload's batch (the reference loops per sample,
    so cost is linear in B).  Returns (Gvcv/s, seconds per step, cores, text)."""
    from oracle import torch_port          # CPU baseline leg: the one allowed use of oracle/
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    f, P, cv, _ = syn.make_inputs(syn.Workload("s", n_samples, w.V, w.C, w.H, w.W, w.G, w.method, w.dtype))
    for _ in range(warmup):
        torch_port.unprojection(f, P, cv, w.method)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        torch_port.unprojection(f, P, cv, w.method)
        times.append(time.perf_counter() - t0)
    units = n_samples * w.G ** 3 * w.C * w.V
    sec = sum(times) / len(times)
    return units / sec / 1e9, sec, cores, "%d of %d samples of %s per step, torch CPU path, %d threads" % (
        n_samples, w.B, w.name, cores)


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    val, sec, cores, sample = cpu_reference_sample(w, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": describe(w), "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def describe(w):
    return "%s: B%d V%d C%d %dx%d -> %d^3, %s fusion, %s features" % (
        w.name, w.B, w.V, w.C, w.H, w.W, w.G, w.method, w.dtype)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="cfg2", choices=sorted(syn.CONFIGS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    w = syn.CONFIGS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
        return
    args.warmup = max(args.warmup, 3)

    from multiviewhmr_b200 import aggregation as agg
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    # rotating input sets: more input bytes in flight than the 126 MB L2 holds
    in_bytes = w.B * w.V * w.C * w.H * w.W * (2 if w.dtype == "bf16" else 4) + w.B * w.G ** 3 * 12
    n_sets = max(2, -(-2 * L2_BYTES // in_bytes))
    host_sets, dev_sets = [], []
    for i in range(n_sets):
        f, P, cv, _ = syn.make_inputs(w, seed=1234 + 97 * rank + i)
        if w.dtype == "bf16":
            f = f.bfloat16()
        host = tuple(t.pin_memory() for t in (f, P, cv))
        host_sets.append(host)
        dev_sets.append(tuple(t.to(dev) for t in host))
    outs = [torch.empty((w.B, w.C, w.G, w.G, w.G), dtype=torch.float32, device=dev) for _ in range(2)]
    stream = torch.cuda.current_stream(dev)

    # ---- device-resident timing: K steps of (pack + fused kernel) ------------------
    kern_ev = []

    def step(i, timed):
        f, P, cv = dev_sets[i % n_sets]
        packed = agg.pack_features(f)                                   # launch 1
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        agg.unprojection(f, P, cv, w.method, out=outs[i % 2], packed=packed)   # launch 2
        if timed:
            e1.record(stream)
            kern_ev.append((e0, e1))

    for i in range(args.warmup):
        step(i, False)
    sampler = ClockSampler(dev)
    barrier()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with sampler:
        t_start.record(stream)
        for i in range(args.steps):
            step(i, True)
        t_end.record(stream)
        barrier()
    elapsed_ms = t_start.elapsed_time(t_end)
    kernel_ms = sum(a.elapsed_time(b) for a, b in kern_ev) / len(kern_ev)
    launches = 2 * args.steps

    # ---- informative extra: the same step when the producer hands over channels-last maps, which the
    # kernel gathers in place (MVHMR_LAYOUT_NHWC: no pack_kernel).  Not the headline: the reference's
    # backbone emits NCHW, which is what `value` is measured on.
    cl_sets = [f.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3) for f, _, _ in dev_sets]
    cl_ok = agg._is_channels_last(cl_sets[0])
    cl_ms = None
    if cl_ok:
        for i in range(args.warmup):
            agg.unprojection(cl_sets[i % n_sets], dev_sets[i % n_sets][1], dev_sets[i % n_sets][2], w.method, out=outs[i % 2])
        barrier()
        c_start, c_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c_start.record(stream)
        for i in range(args.steps):
            agg.unprojection(cl_sets[i % n_sets], dev_sets[i % n_sets][1], dev_sets[i % n_sets][2], w.method, out=outs[i % 2])
        c_end.record(stream)
        barrier()
        cl_ms = c_start.elapsed_time(c_end) / args.steps
    del cl_sets

    # ---- end to end: pinned host inputs -> public API -> metric back on the host ------
    e2e_steps = max(5, min(args.steps, 30))
    metric_host = torch.empty((w.B, w.C, 3), dtype=torch.float32).pin_memory()
    d2h = metric_host.numel() * 4
    # Host inputs of a step, as in VolumeGenerator.forward (models/aggregation.py:119-193): feature maps,
    # projection matrices and the per-sample cuboid centre / rotation; the coordinate volume itself is
    # built on the device (the reference builds it there too, :150-187).
    rots_np = np.stack([np.eye(3, dtype=np.float32)] * w.B)
    centers_sets = [syn.make_inputs(syn.Workload("c", w.B, 1, 1, 1, 1, 2), seed=1234 + 97 * rank + i)[3].numpy()
                    for i in range(n_sets)]
    h2d = sum(t.numel() * t.element_size() for t in host_sets[0][:2]) + w.B * 12 * 4

    # Double-buffered: the host->device copies of step i+1 run on a copy stream while step i
    # computes; every copy and every kernel of all e2e steps is inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    dev_in = [tuple(torch.empty_like(t, device=dev) for t in host_sets[0][:2]) for _ in range(2)]
    grid_in = [None, None]
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def stage(i):
        slot = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])            # the slot's previous user is done
            for d, h in zip(dev_in[slot], host_sets[i % n_sets][:2]):
                d.copy_(h, non_blocking=True)
            grid_in[slot] = agg.build_coord_volumes(centers_sets[i % n_sets], rots_np, w.G, w.cuboid_side, dev)
            copied[slot].record(copy_stream)

    def e2e_run(n):
        for ev in consumed:
            ev.record(stream)
        stage(0)
        for i in range(n):
            slot = i % 2
            if i + 1 < n:
                stage(i + 1)
            stream.wait_event(copied[slot])
            f, P = dev_in[slot]
            cv = grid_in[slot]
            cv.record_stream(stream)
            vol = agg.unprojection(f, P, cv, w.method, out=outs[slot])   # pack + fused kernel
            joints = agg.soft_argmax_3d(vol, cv)                         # the step's metric: (B,C,3) expectations
            metric_host.copy_(joints, non_blocking=True)
            consumed[slot].record(stream)

    e2e_run(3)
    barrier()
    # three repetitions of e2e_steps steps, median reported: one host hiccup (page faults of a fresh box,
    # a pinned-pool growth) otherwise decides the whole number
    e2e_reps = []
    with sampler:
        for _ in range(3):
            e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e_start.record(stream)
            e2e_run(e2e_steps)
            e_end.record(stream)
            barrier()
            e2e_reps.append(e_start.elapsed_time(e_end))
    e2e_ms = sorted(e2e_reps)[1]

    if world > 1:
        t = torch.tensor([elapsed_ms, e2e_ms, kernel_ms, cl_ms or 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms, kernel_ms, cl_max = (float(x) for x in t.cpu())
        cl_ms = cl_max if cl_ok else None

    if rank == 0:
        units = w.vcv * world
        ms_per_step = elapsed_ms / args.steps
        value = units / (ms_per_step * 1e-3) / 1e9
        peak, peak_src = measured_peak()
        alg = w.algorithmic_bytes()
        achieved = alg / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if w.dtype == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": describe(w), "per_gpu_batch": w.B, "parallelism": "batch-sharded x%d, no collective" % world,
                       "l2": "%d rotating input sets (%.0f MB > 126 MB L2), outputs double-buffered" % (n_sets, n_sets * in_bytes / 1e6),
                       "step": "pack_kernel + unproject_kernel", "tile": os.environ.get("MVHMR_TILE", "auto")},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": recorded_traffic(w.name), "kernel": "unproject_kernel",
                         "kernel_ms": kernel_ms, "algorithmic_bytes": alg, "peak_source": peak_src,
                         "step_frac": (alg / (ms_per_step * 1e-3) / 1e9) / peak},
            "e2e": {"value": units / (e2e_ms / e2e_steps * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "repetitions_ms_per_step": [t / e2e_steps for t in e2e_reps],
                    "path": "pinned host features+proj+centres -> (copy stream, double-buffered) build_coord_volumes() -> unprojection() -> soft_argmax_3d() -> host"},
            "gpu_launches": launches * world,
            "clocks": sampler.summary(),
        }
        if cl_ms:
            line["extras"] = {"channels_last_in_place": {
                "ms_per_step": cl_ms, "value": units / (cl_ms * 1e-3) / 1e9, "unit": UNIT,
                "note": "same workload, feature maps handed over as (B,V,H,W,C): unproject_kernel only, no pack_kernel"}}
        if world == 1 and not args.no_cpu_baseline:
            val, sec, cores, sample = cpu_reference_sample(w, steps=2, warmup=1)
            line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
