#!/usr/bin/env python
"""Benchmark of the volumetric-aggregation hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg5] [--impl reference]

Workload.  BASELINE.json quotes its metric "at 1/2/4/8 B200" on configs[4], the
scaling sweep (cfg5: B64 V8 C32 96x96 -> 80^3, softmax fusion, fp32), which fits
one GPU — so cfg5 is the headline at every N.  A step = one pass of the hot path
(feature packing + fused unproject+aggregate) over this rank's share of the
batch.  For N>1 the driver launches this file under torchrun, one rank per GPU;
the B*Gx x-planes of the batch are cut with `sharding.shard_windows` (whole
samples when N divides B — the reference's DDP batch split, train.py:166-168 —
x-slabs otherwise), no collective in the data path, STRONG scaling: the total
work is fixed and `value` is all ranks' units over the slowest rank's device time.

Rank 0 prints ONE JSON line.
  value      device-resident (inputs already in HBM), CUDA events, max over ranks
  roofline   the fused kernel alone, CUDA events around its launches in the timed loop
  e2e        the same step through the public Python API with pinned HOST inputs: H2D of the
             feature maps / projections / coordinate volumes, pack + fused kernel, and the
             D2H of the FULL aggregated volume (B,C,G,G,G) into pinned host memory, all
             inside the timed region (three streams, double-buffered).  `e2e.consumer_on_gpu`
             is the variant whose consumer stays on the GPU (soft-argmax -> (B,C,3) to host).
  configs    (N=1 only) every BASELINE config cfg1..cfg5 device-resident: ms, Gvcv/s and the
             fraction of the HBM roofline of the fused kernel; the small configs also replayed from
             CUDA graphs (`cuda_graph`); cfg3 also its soft-argmax, the ONE-kernel fused
             unproject+aggregate+soft-argmax path (`fused_soft_argmax`: volume stored / joints only)
             and the reduced-precision texture path (`fast_path`)
  extras.backward (N=1 only) the gradient kernels at cfg2 and cfg4 with their roofline
  extras.all_gather_volume (N>1 only) the optional NCCL all-gather of the sharded volume, timed
             separately — it is not part of `value`

`--impl reference` times the reference's own CPU path (oracle/torch_port.py: the same
ATen calls in the same order, all host threads) on a bounded sample of the same
workload.  That is the only place besides `cpu_baseline` where this file executes
anything under oracle/.
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

from multiviewhmr_b200 import sharding, synthetic as syn  # noqa: E402

METRIC = "Gvoxel-ch-views/s fused unproject+aggregate"
UNIT = "Gvcv/s"
L2_BYTES = 126 * 1024 * 1024
HEADLINE = "cfg5"


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
    """SM clock + throttle reasons sampled while the timed region runs: once when it starts, every
    50 ms, once when it ends.  (NVML queries share a driver lock with kernel launches: polling every
    2 ms, as round 1 did, starved the GPU between the two launches of an 8 ms cfg5 step.)"""
    PERIOD = 0.05
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, dev):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            if os.environ.get("MVHMR_BENCH_NO_SAMPLER"):
                raise RuntimeError("sampler disabled")
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

    def _sample(self):
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self.REASONS.items():
                if bits & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _run(self):
        while not self._stop.wait(self.PERIOD):
            self._sample()
        self._sample()

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


def sub_workload(w, B, name=None):
    return syn.Workload(name or w.name, B, w.V, w.C, w.H, w.W, w.G, w.method, w.dtype, w.joints, w.cuboid_side)


def cpu_reference_sample(w, steps, warmup, n_samples=2):
    """Time the torch port of the reference on the host cores: the first
    `n_samples` samples of the workload's batch (the reference loops per sample,
    so cost is linear in B).  Returns (Gvcv/s, seconds per step, cores, text)."""
    from oracle import torch_port          # CPU baseline leg: the one allowed use of oracle/
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_samples = min(n_samples, w.B)
    f, P, cv, _ = syn.make_inputs(sub_workload(w, n_samples, "s"))
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


def config_of(w, world):
    """Identical in both arms (the driver compares the two lines' `config`)."""
    return {"workload": describe(w), "global_batch": w.B,
            "parallelism": "shard_windows over %d rank(s): batch split%s, no collective in the data path"
                           % (world, "" if w.B % world == 0 else " + x-slabs")}


def describe(w):
    return "%s: B%d V%d C%d %dx%d -> %d^3, %s fusion, %s features" % (
        w.name, w.B, w.V, w.C, w.H, w.W, w.G, w.method, w.dtype)


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    val, sec, cores, sample = cpu_reference_sample(w, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(w, args.gpus),
        "notes": {"sample": sample, "arm": "the reference's torch CPU path on rank 0's host cores; no GPU work"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class RankShare:
    """This rank's windows of workload `w` and synthetic inputs for exactly the samples they touch."""

    def __init__(self, w, rank, world, dev, n_sets_min=1):
        self.w, self.dev = w, dev
        self.windows = sharding.shard_windows(w.B, w.G, rank, world)
        self.b_lo = min([x.b0 for x in self.windows], default=0)
        self.b_hi = max([x.b1 for x in self.windows], default=0)
        self.nb = self.b_hi - self.b_lo
        self.units = sum(x.units() for x in self.windows) * w.G * w.G * w.C * w.V     # voxel-channel-views of this rank
        e = 2 if w.dtype == "bf16" else 4
        self.in_bytes = self.nb * (w.V * w.C * w.H * w.W * e + w.G ** 3 * 12 + w.V * 48)
        self.out_bytes = self.nb * w.C * w.G ** 3 * 4
        # rotating input sets: more input bytes in flight than the 126 MB L2 holds
        self.n_sets = max(n_sets_min, min(8, -(-2 * L2_BYTES // max(self.in_bytes, 1)))) if self.nb else 0
        lw = sub_workload(w, self.nb)
        self.host_sets, self.dev_sets = [], []
        for i in range(self.n_sets):
            f, P, cv, centers = syn.make_inputs(lw, seed=1234 + 1000 * i + self.b_lo, b_offset=self.b_lo)
            if w.dtype == "bf16":
                f = f.bfloat16()
            self.host_sets.append((f, P, centers.numpy()))
            self.dev_sets.append(tuple(t.to(dev) for t in (f, P, cv)))
        self.outs = [torch.empty((self.nb, w.C, w.G, w.G, w.G), dtype=torch.float32, device=dev) for _ in range(2)]
        # the packed-plane workspace of a step, allocated once (two, so that the pack of step i+1 never waits
        # for a reader of step i in the end-to-end pipeline)
        self.packs = [None, None]

    def pack(self, agg, f, i):
        if self.packs[i % 2] is None:
            self.packs[i % 2] = agg.pack_features(f)
            return self.packs[i % 2]
        return agg.pack_features(f, out=self.packs[i % 2])

    def local_windows(self):
        gy = gz = self.w.G
        for x in self.windows:
            n0, n1 = x.voxels(gy, gz)
            yield (x.b0 - self.b_lo, x.b1 - self.b_lo, n0, n1)

    def launches_per_step(self):
        return 1 + len(self.windows)            # pack_kernel + one fused kernel per window

    def step(self, agg, i, out=None, inputs=None):
        """pack + fused kernel(s) over this rank's windows; returns the output buffer."""
        f, P, cv = inputs if inputs is not None else self.dev_sets[i % self.n_sets]
        out = self.outs[i % 2] if out is None else out
        packed = self.pack(agg, f, i)
        for win in self.local_windows():
            agg.unprojection(f, P, cv, self.w.method, window=win, out=out, packed=packed)
        return out


def time_device_resident(agg, share, steps, warmup, stream, barrier, sampler=None):
    """K steps of (pack + fused kernel); returns (ms/step over the whole loop, fused-kernel ms/step)."""
    if share.nb == 0:
        barrier(); barrier()
        return 0.0, 0.0
    kern_ev = []
    for i in range(warmup):
        share.step(agg, i)
    barrier()

    def loop():
        for i in range(steps):
            f, P, cv = share.dev_sets[i % share.n_sets]
            out = share.outs[i % 2]
            packed = share.pack(agg, f, i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for win in share.local_windows():
                agg.unprojection(f, P, cv, share.w.method, window=win, out=out, packed=packed)
            e1.record(stream)
            kern_ev.append((e0, e1))
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler is not None:
        with sampler:
            t0.record(stream); loop(); t1.record(stream)
            barrier()
    else:
        t0.record(stream); loop(); t1.record(stream)
        barrier()
    return t0.elapsed_time(t1) / steps, sum(a.elapsed_time(b) for a, b in kern_ev) / len(kern_ev)


def time_e2e(agg, share, steps, stream, barrier, dev, full_d2h, sampler):
    """pinned host inputs -> H2D -> pack + fused kernel -> D2H of the result, three streams,
    double-buffered; every copy and kernel of all steps is inside the timed region.
    Returns (median ms per step of 3 repetitions, the repetitions, h2d bytes, d2h bytes)."""
    w = share.w
    if share.nb == 0:
        for _ in range(5):
            barrier()
        return 0.0, [0.0, 0.0, 0.0], 0, 0
    # Host inputs of a step, as in VolumeGenerator.forward (models/aggregation.py:119-193): feature maps,
    # projection matrices and the per-sample cuboid centre / rotation; the coordinate volume itself is
    # built on the device (the reference builds it there too, :150-187).
    host_sets = [tuple(t.pin_memory() for t in hs[:2]) for hs in share.host_sets]
    centers_sets = [hs[2] for hs in share.host_sets]
    rots_np = np.stack([np.eye(3, dtype=np.float32)] * share.nb)
    h2d = sum(t.numel() * t.element_size() for t in host_sets[0]) + share.nb * 12 * 4
    if full_d2h:
        host_out = [torch.empty((share.nb, w.C, w.G, w.G, w.G), dtype=torch.float32).pin_memory() for _ in range(2)]
    else:
        host_out = [torch.empty((share.nb, w.C, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
    d2h = host_out[0].numel() * 4
    copy_in, copy_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    dev_in = [tuple(torch.empty_like(t, device=dev) for t in host_sets[0]) for _ in range(2)]
    grid_in = [None, None]
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    computed = [torch.cuda.Event() for _ in range(2)]
    drained = [torch.cuda.Event() for _ in range(2)]

    def stage(i):
        slot = i % 2
        with torch.cuda.stream(copy_in):
            copy_in.wait_event(consumed[slot])                # the slot's previous user is done
            for d, h in zip(dev_in[slot], host_sets[i % len(host_sets)]):
                d.copy_(h, non_blocking=True)
            grid_in[slot] = agg.build_coord_volumes(centers_sets[i % len(host_sets)], rots_np, w.G, w.cuboid_side, dev)
            copied[slot].record(copy_in)

    def run(n):
        for ev in consumed + drained:
            ev.record(stream)
        stage(0)
        for i in range(n):
            slot = i % 2
            if i + 1 < n:
                stage(i + 1)
            stream.wait_event(copied[slot])
            stream.wait_event(drained[slot])                  # the output buffer's previous D2H is done
            cv = grid_in[slot]
            cv.record_stream(stream)
            vol = share.step(agg, i, out=share.outs[slot], inputs=dev_in[slot] + (cv,))
            if full_d2h:
                computed[slot].record(stream)
                consumed[slot].record(stream)
                with torch.cuda.stream(copy_out):
                    copy_out.wait_event(computed[slot])
                    host_out[slot].copy_(vol, non_blocking=True)
                    drained[slot].record(copy_out)
            else:
                joints = agg.soft_argmax_3d(vol, cv)                  # the consumer stays on the GPU
                host_out[slot].copy_(joints, non_blocking=True)
                consumed[slot].record(stream)
                drained[slot].record(stream)
        stream.wait_stream(copy_out)

    run(2)
    barrier()
    reps = []
    with sampler:
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            run(steps)
            e1.record(stream)
            barrier()
            reps.append(e0.elapsed_time(e1) / steps)
    del host_out, dev_in, host_sets
    return sorted(reps)[1], reps, h2d, d2h


def graph_replay_ms(share, agg, steps, stream, barrier, body=None):
    """The step (pack + fused kernel, or `body`) captured ONCE per rotating input set into a CUDA graph and
    replayed: the library's calls are capturable (no allocation, no host sync), so the small configs —
    whose eager loop is bound by ~0.1 ms of Python / ctypes per call, not by the GPU — are timed as the
    device executes them.  Returns ms per step, or None if capture is refused."""
    try:
        graphs = []
        for k in range(share.n_sets):
            f, P, cv = share.dev_sets[k]
            out = share.outs[k % 2]
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                if body is None:
                    agg.unprojection(f, P, cv, share.w.method, out=out, packed=share.pack(agg, f, k))
                else:
                    body(f, P, cv, out)
            graphs.append(g)
        for g in graphs:
            g.replay()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            graphs[i % len(graphs)].replay()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1) / steps
        del graphs
        return ms
    except Exception as exc:                      # noqa: BLE001 — a refused capture must not cost the bench line
        sys.stderr.write("graph capture refused: %r\n" % (exc,))
        torch.cuda.synchronize()
        return None


def per_config_block(agg, dev, stream, barrier, peak, headline_name, headline_entry):
    """cfg1..cfg5 device-resident on one GPU: pack + fused kernel (cfg3: + soft-argmax)."""
    block = {}
    for name in sorted(syn.CONFIGS):
        if name == headline_name:
            block[name] = headline_entry
            continue
        w = syn.CONFIGS[name]
        share = RankShare(w, 0, 1, dev)
        steps = 60 if w.vcv < 1e9 else 20
        ms, kms = time_device_resident(agg, share, steps, 3, stream, barrier)
        alg = w.algorithmic_bytes()
        entry = {"workload": describe(w), "ms_per_step": ms, "kernel_ms": kms, "value": w.vcv / (ms * 1e-3) / 1e9, "unit": UNIT,
                 "algorithmic_bytes": alg, "roofline_frac": alg / (kms * 1e-3) / 1e9 / peak,
                 "step_frac": alg / (ms * 1e-3) / 1e9 / peak, "steps": steps}
        if w.vcv < 2e9:
            gms = graph_replay_ms(share, agg, steps, stream, barrier)
            if gms is not None:
                entry["cuda_graph"] = {"ms_per_step": gms, "value": w.vcv / (gms * 1e-3) / 1e9, "unit": UNIT,
                                       "step_frac": alg / (gms * 1e-3) / 1e9 / peak,
                                       "note": "the same step replayed from a CUDA graph (no per-call host overhead)"}
        if w.joints:
            f, P, cv = share.dev_sets[0]
            vol = share.step(agg, 0)
            for _ in range(3):
                agg.soft_argmax_3d(vol[:, :w.joints], cv)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(steps):
                agg.soft_argmax_3d(share.outs[i % 2][:, :w.joints], share.dev_sets[i % share.n_sets][2])
            e1.record(stream)
            barrier()
            sa_ms = e0.elapsed_time(e1) / steps
            entry["soft_argmax"] = {"joints": w.joints, "ms": sa_ms, "algorithmic_bytes": w.soft_argmax_bytes(),
                                    "roofline_frac": w.soft_argmax_bytes() / (sa_ms * 1e-3) / 1e9 / peak}
            # the eager loop above is bound by the host (~50 us of Python per call): the same two launches from a graph
            gms = graph_replay_ms(share, agg, steps, stream, barrier,
                                  body=lambda f, P, cv, out: agg.soft_argmax_3d(out[:, :w.joints], cv))
            if gms is not None:
                entry["soft_argmax"]["cuda_graph_ms"] = gms
                entry["soft_argmax"]["cuda_graph_roofline_frac"] = w.soft_argmax_bytes() / (gms * 1e-3) / 1e9 / peak
            # BASELINE.json's target path as ONE kernel: unproject + aggregate + soft-argmax
            # (mvhmr_unproject_aggregate_softargmax), with and without the volume store
            fused = {"two_kernel_step_ms": ms + sa_ms}
            two = agg.soft_argmax_3d(vol[:, :w.joints], cv)
            for label, store in (("volume_stored", True), ("joints_only", False)):
                for i in range(3):
                    f, P, cv = share.dev_sets[i % share.n_sets]
                    j = agg.unprojection_soft_argmax(f, P, cv, w.joints, w.method, store_volume=store)[1]
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for i in range(steps):
                    f, P, cv = share.dev_sets[i % share.n_sets]
                    agg.unprojection_soft_argmax(f, P, cv, w.joints, w.method, store_volume=store)
                e1.record(stream)
                barrier()
                fms = e0.elapsed_time(e1) / steps
                bytes_ = alg if store else alg - w.B * w.C * w.G ** 3 * 4
                fused[label] = {"ms_per_step": fms, "value": w.vcv / (fms * 1e-3) / 1e9, "unit": UNIT,
                                "algorithmic_bytes": bytes_, "step_frac": bytes_ / (fms * 1e-3) / 1e9 / peak}
                gms = graph_replay_ms(share, agg, steps, stream, barrier,
                                      body=lambda f, P, cv, out, store=store: agg.unprojection_soft_argmax(
                                          f, P, cv, w.joints, w.method, store_volume=store))
                if gms is not None:
                    fused[label]["cuda_graph_ms_per_step"] = gms
            gms = graph_replay_ms(share, agg, steps, stream, barrier,
                                  body=lambda f, P, cv, out: agg.soft_argmax_3d(
                                      agg.unprojection(f, P, cv, w.method, out=out, packed=agg.pack_features(f))[:, :w.joints], cv))
            if gms is not None:
                fused["two_kernel_cuda_graph_ms_per_step"] = gms
            f, P, cv = share.dev_sets[0]
            j = agg.unprojection_soft_argmax(f, P, cv, w.joints, w.method, store_volume=False)[1]
            fused["max_abs_diff_vs_two_kernel_mm"] = float((j - two).abs().max())
            fused["note"] = ("pack + ONE fused kernel + record merge; joints_only never writes or re-reads the volume "
                             "(its roofline numerator drops the B*C*G^3*4 output bytes)")
            entry["fused_soft_argmax"] = fused
        if w.dtype == "bf16":
            # the opt-in reduced-precision path (texture units, fp16 maps): inside the bf16 tolerance (1e-2)
            f, P, cv = share.dev_sets[0]
            exact = share.step(agg, 0).clone()
            fast = agg.unprojection(f, P, cv, w.method, precision="fast")
            dev_rel = float((fast.double() - exact.double()).norm() / exact.double().norm())
            del exact, fast
            for i in range(3):
                agg.unprojection(f, P, cv, w.method, out=share.outs[i % 2], precision="fast")
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(steps):
                f, P, cv = share.dev_sets[i % share.n_sets]
                agg.unprojection(f, P, cv, w.method, out=share.outs[i % 2], precision="fast")
            e1.record(stream)
            barrier()
            fms = e0.elapsed_time(e1) / steps
            entry["fast_path"] = {"ms_per_step": fms, "value": w.vcv / (fms * 1e-3) / 1e9, "unit": UNIT,
                                  "step_frac": alg / (fms * 1e-3) / 1e9 / peak, "rel_l2_vs_exact": dev_rel,
                                  "tolerance": 1e-2, "speedup_vs_exact_step": ms / fms,
                                  "note": "precision='fast': tex_pack_kernel + unproject_tex_kernel (hardware bilinear filtering of fp16 texels)"}
        block[name] = entry
        del share
        torch.cuda.empty_cache()
    return block


def output_format_block(agg, dev, stream, barrier):
    """Consumer-side output formats of the fused kernel at cfg2: device time over pre-packed planes."""
    w = syn.CONFIGS["cfg2"]
    f, P, cv, _ = syn.make_inputs(w)
    fd, Pd, cvd = (t.to(dev) for t in (f, P, cv))
    packed = agg.pack_features(fd)
    block = {}
    for fmt in ("ncdhw", "channels_last_3d", "max_pool2"):
        out = agg.unprojection(fd, Pd, cvd, w.method, packed=packed, output=fmt)
        for _ in range(3):
            agg.unprojection(fd, Pd, cvd, w.method, packed=packed, out=out, output=fmt)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(30):
            agg.unprojection(fd, Pd, cvd, w.method, packed=packed, out=out, output=fmt)
        e1.record(stream)
        barrier()
        block[fmt] = {"kernel_ms": e0.elapsed_time(e1) / 30, "output_bytes": out.numel() * 4}
    return block


def backward_block(dev, stream, barrier, peak):
    """Gradient w.r.t. the feature maps (training goes through it, train.py:110): sum / max / softmax at
    cfg2 and cfg4.  Algorithmic bytes: grad_out read once, features read once (max / softmax re-sample),
    gradient written once."""
    from multiviewhmr_b200 import autograd
    block = {}
    for name in ("cfg2", "cfg4"):
        w = syn.CONFIGS[name]
        f, P, cv, _ = syn.make_inputs(w)
        fd, Pd, cvd = (t.to(dev) for t in (f, P, cv))
        g = torch.randn((w.B, w.C, w.G, w.G, w.G), device=dev)
        alg = w.B * w.C * w.G ** 3 * 4 + 2 * w.B * w.V * w.C * w.H * w.W * 4 + w.B * w.G ** 3 * 12
        steps = 20 if name == "cfg2" else 6
        for method in ("sum", "max", "softmax"):
            for _ in range(2):
                autograd.unprojection_backward(g, fd, Pd, cvd, method)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                autograd.unprojection_backward(g, fd, Pd, cvd, method)
            e1.record(stream)
            barrier()
            ms = e0.elapsed_time(e1) / steps
            block["%s_%s" % (name, method)] = {"ms": ms, "algorithmic_bytes": alg,
                                               "roofline_frac": alg / (ms * 1e-3) / 1e9 / peak, "steps": steps}
        del fd, cvd, g
        torch.cuda.empty_cache()
    return block


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=HEADLINE, choices=sorted(syn.CONFIGS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the per-config and backward blocks (N=1)")
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

    stream = torch.cuda.current_stream(dev)
    peak, peak_src = measured_peak()
    share = RankShare(w, rank, world, dev)
    sampler = ClockSampler(dev)

    # ---- device-resident timing: K steps of (pack + fused kernel) over this rank's windows ----
    ms_per_step, kernel_ms = time_device_resident(agg, share, args.steps, args.warmup, stream, barrier, sampler)

    # ---- end to end: pinned host inputs -> public API -> result back in pinned host memory ----
    e2e_steps = max(4, min(args.steps, 8 if share.out_bytes > (1 << 30) else 20))
    e2e_ms, e2e_reps, h2d, d2h = time_e2e(agg, share, e2e_steps, stream, barrier, dev, True, sampler)
    gpu_ms, gpu_reps, _, gpu_d2h = time_e2e(agg, share, e2e_steps, stream, barrier, dev, False, sampler)

    # ---- optional collective, reported separately: every rank ends with the full volume ----
    gather = None
    if world > 1 and w.B % world == 0 and not args.no_extras:
        per = w.B // world
        full = torch.empty((w.B, w.C, w.G, w.G, w.G), dtype=torch.float32, device=dev)
        full[rank * per:(rank + 1) * per].copy_(share.outs[0])
        sharding.all_gather_volume(full, w.B, w.G, world)           # warm-up (NCCL channel setup)
        barrier()
        reps = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            sharding.all_gather_volume(full, w.B, w.G, world)
            e1.record(stream)
            barrier()
            reps.append(e0.elapsed_time(e1))
        tg = torch.tensor([sorted(reps)[1]], device=dev)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        nbytes = full.numel() * 4
        gather = {"ms": float(tg[0]), "bytes_total": nbytes, "busbw_gbs": nbytes * (world - 1) / world / (float(tg[0]) * 1e-3) / 1e9,
                  "note": "sharding.all_gather_volume: ONE ncclAllGather straight into the full (B,C,G,G,G) volume over NVLink; "
                          "NOT part of `value` — data-parallel consumers keep the volume sharded"}
        del full
        torch.cuda.empty_cache()

    if world > 1:
        t = torch.tensor([ms_per_step, e2e_ms, kernel_ms, gpu_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step, e2e_ms, kernel_ms, gpu_ms = (float(x) for x in t.cpu())
        cnt = torch.tensor([float(share.units), float(share.launches_per_step())], device=dev, dtype=torch.float64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        total_units, launches_per_step = float(cnt[0]), int(cnt[1])
    else:
        total_units, launches_per_step = float(share.units), share.launches_per_step()
    assert total_units == float(w.vcv), (total_units, w.vcv)

    if rank == 0:
        value = total_units / (ms_per_step * 1e-3) / 1e9
        # roofline of the fused kernel: the slowest rank's launches cover 1/world of the algorithmic bytes
        alg = w.algorithmic_bytes()
        alg_rank = alg / world
        achieved = alg_rank / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16" if w.dtype == "bf16" else "f32", "data": "synthetic",
            "config": config_of(w, world),
            "notes": {"per_gpu_batch": w.B / world,
                      "l2": "%d rotating input set(s) of %.0f MB per rank + %.0f MB of output per step (L2 = 126 MB), outputs double-buffered"
                            % (share.n_sets, share.in_bytes / 1e6, share.out_bytes / 1e6),
                      "step": "pack_kernel + fused unproject/aggregate kernel", "path": os.environ.get("MVHMR_PATH", "auto")},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": recorded_traffic(w.name), "kernel": "fused unproject+aggregate kernel",
                         "kernel_ms": kernel_ms, "algorithmic_bytes": alg_rank, "peak_source": peak_src,
                         "step_frac": (alg_rank / (ms_per_step * 1e-3) / 1e9) / peak,
                         "nominal_8tbs_frac": achieved / 8000.0},
            "e2e": {"value": total_units / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": e2e_ms,
                    "repetitions_ms_per_step": e2e_reps,
                    "h2d_gbs_per_rank": h2d / (e2e_ms * 1e-3) / 1e9, "d2h_gbs_per_rank": d2h / (e2e_ms * 1e-3) / 1e9,
                    "path": "pinned host features+proj+centres -> copy-in stream (+ build_coord_volumes()) -> pack_features() + unprojection() -> "
                            "copy-out stream -> the full (B,C,G,G,G) volume in pinned host memory (bytes are per rank)",
                    "consumer_on_gpu": {"value": total_units / (gpu_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": gpu_ms,
                                        "d2h_bytes_per_step": gpu_d2h,
                                        "path": "same inputs; the volume stays on the GPU, soft_argmax_3d() -> (B,C,3) to host"}},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": sampler.summary(),
        }
        if gather is not None:
            line["extras"] = {"all_gather_volume": gather}
        if world == 1 and not args.no_extras:
            head = {"workload": describe(w), "ms_per_step": ms_per_step, "kernel_ms": kernel_ms, "value": value, "unit": UNIT,
                    "algorithmic_bytes": alg, "roofline_frac": achieved / peak,
                    "step_frac": line["roofline"]["step_frac"], "steps": args.steps}
            del share
            torch.cuda.empty_cache()
            line["configs"] = per_config_block(agg, dev, stream, barrier, peak, w.name, head)
            line["extras"] = {"backward": backward_block(dev, stream, barrier, peak),
                              "output_formats_cfg2": output_format_block(agg, dev, stream, barrier)}
        if world == 1 and not args.no_cpu_baseline:
            val, sec, cores, sample = cpu_reference_sample(w, steps=3, warmup=1)
            line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
