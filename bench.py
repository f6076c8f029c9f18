#!/usr/bin/env python
"""bench.py - pairs/sec of PWC-Net's warp + cost-volume hot path on B200 (BASELINE.json metric).

One "step" = one pass of the pyramid loop's hot path (tfoptflow/model_pwcnet.py:1550-1567: 5 cost volumes,
4 warps) over one batch of synthetic feature pyramids.  Default workload = BASELINE.json configs[2]:
Sintel 436x1024 padded to 448x1024, batch 32 per GPU, search_range 4, fp32, forward.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 0|1|2|4] [--mode fwd|fwdbwd|allreduce]
                  [--dtype fp32|fp16] [--scaling weak|strong] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...       (one rank per GPU)

  --config 2 (default)  448x1024, batch 32, r=4, forward (fp32 unless --dtype fp16)       BASELINE configs[2]
  --config 0            384x512, batch 1, r=4, fp32 forward                                BASELINE configs[0] (operator part)
  --config 1            384x512, batch 8, six levels all warped, fp32 forward + backward   BASELINE configs[1]
  --config 3            pwcnet-sm-6-2 training step, 384x448 crops, batch 8 per GPU, mixed precision, data-parallel
                        with one NCCL all-reduce of the gradients (needs the PyTorch PWC-Net of model_pwcnet.py)  configs[3]
  --config 4            1088x1920, batch 16, r=8 (289 channels), fp16 forward              BASELINE configs[4]
  --mode allreduce      the training configuration's one exchange step (multi_gpus.average_gradients over NCCL)
  --mode infer          the whole PWC-Net forward of model_pwcnet.py at the preset's shape (use with --config 0)

Prints ONE JSON line (rank 0).  `value` is device-resident throughput, `e2e` goes through the host-buffer path with
H2D/D2H inside the timed region, `roofline` is the dominant kernel (pyramid level 2; its backward in fwdbwd mode)
against the measured HBM copy bandwidth - with the CUDA-core FMA bound next to it, because this path needs
~70 TFLOP/s of fp32 FMA to run at HBM speed - and `cpu_baseline` is the op-for-op CPU port of the reference graph
(TensorFlow 1.10 is not installable) on the box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "pairs/sec for PWC-Net costvol+warp at 436x1024 (padded 448x1024)"
UNIT = "pairs/s"
FMA_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # CUDA-core fp32 FMA peak of a B200: 74.4

PRESETS = {
    0: dict(height=384, width=512, batch=1, search_range=4, dtype="fp32", mode="fwd", levels=(6, 5, 4, 3, 2), top_flow=False),
    1: dict(height=384, width=512, batch=8, search_range=4, dtype="fp32", mode="fwdbwd", levels=(6, 5, 4, 3, 2, 1), top_flow=True),
    2: dict(height=448, width=1024, batch=32, search_range=4, dtype="fp32", mode="fwd", levels=(6, 5, 4, 3, 2), top_flow=False),
    3: dict(height=384, width=448, batch=8, search_range=4, dtype="fp16", mode="train", levels=(6, 5, 4, 3, 2), top_flow=False),
    4: dict(height=1088, width=1920, batch=16, search_range=8, dtype="fp16", mode="fwd", levels=(6, 5, 4, 3, 2), top_flow=False),
}
OPTION_DEFAULTS = {"fwd_impl": 0, "bwd_impl": 0, "stream_ctas": 0, "stream_prefetch": 4, "stream_debug": 0, "bwd_debug": 0, "bwd_prefetch": 1, "mma_tile": 1,
                   "overlap_levels": 1, "bwd_tiled": 1, "side_reverse": 1, "tile_small": 1, "side_sms": None}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(PRESETS))
    ap.add_argument("--mode", default=None, choices=["fwd", "fwdbwd", "allreduce", "train", "infer"])
    ap.add_argument("--dtype", default=None, choices=["fp32", "fp16"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch pairs on every GPU; strong: --batch pairs in total, batch/N per GPU (configs[2])")
    ap.add_argument("--batch", type=int, default=None, help="pairs per GPU per step (weak) / in total (strong)")
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--search-range", type=int, default=None)
    ap.add_argument("--flow", default="smooth", choices=["smooth", "stress", "zero"])
    ap.add_argument("--fwd-impl", type=int, default=0, help="library option fwd_impl (0 = auto)")
    ap.add_argument("--opt", action="append", default=[], help="library option key=value (experiments)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the oracle spot check of the timed outputs")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 20)")
    ap.add_argument("--per-level", action="store_true", help="one call per level, per-level timings on stderr")
    ap.add_argument("--params", default="sm", choices=["sm", "lg"], help="allreduce mode: gradient size (README.md:55)")
    a = ap.parse_args()
    pre = PRESETS[a.config]
    for k in ("height", "width", "batch", "dtype", "mode"):
        if getattr(a, k) is None:
            setattr(a, k, pre[k])
    if a.search_range is None:
        a.search_range = pre["search_range"]
    a.levels, a.top_flow = pre["levels"], pre["top_flow"]
    return a


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def roofline_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler:
    """Samples SM clock and throttle reasons with NVML every few ms while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.004):
        self.index, self.period, self.samples, self.reasons = index, period_s, [], set()
        self.max_mhz, self._stop, self._th, self.ok = None, threading.Event(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # NVML missing: report it, do not fail the bench
            self.err = repr(e)

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(get_reasons(self.h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.ok:
            self._th = threading.Thread(target=self._loop, daemon=True)
            self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._th:
            self._th.join(timeout=1.0)

    def summary(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": getattr(self, "err", "nvml unavailable")}
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def np_dtype(name):
    return np.float32 if name == "fp32" else np.float16


def total_batch(a, world):
    return a.batch if a.scaling == "strong" else a.batch * world


def workload_config(a, world):
    """The workload, identically worded by both arms (the driver compares the two `config` objects)."""
    lv = f"L{max(a.levels)}..L{min(a.levels)}"
    what = {"fwd": "forward", "fwdbwd": "forward + backward"}[a.mode]
    return {"workload": f"BASELINE.json configs[{a.config}]: PWC-Net pyramid hot path ({len(a.levels)} cost volumes {lv}, "
                        f"{len(a.levels) - (0 if a.top_flow else 1)} warps) on {a.height}x{a.width}, search_range "
                        f"{a.search_range}, {a.dtype}, {what}",
            "pairs_per_step": total_batch(a, world), "flow": a.flow, "mode": a.mode, "scaling": a.scaling}


def flops_per_pair(a, backward=False):
    from tfoptflow_b200 import synth
    d = (2 * a.search_range + 1) ** 2
    f = 0
    for (l, h, w, c) in synth.level_shapes(a.height, a.width, a.levels):
        f += h * w * 2 * d * c * (2 if backward else 1)
    return f


# ------------------------------------------------------------------------------------------
# CPU arm: the reference graph's op-for-op port (oracle/ref_graph_torch.py) on the host cores
# ------------------------------------------------------------------------------------------
def cpu_port_time(a, batch, reps, budget_s):
    import torch
    from oracle import ref_graph_torch as T
    from tfoptflow_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    bwd = a.mode == "fwdbwd"
    lv = synth.make_pyramid(batch, a.height, a.width, np.float32, levels=a.levels, flow=a.flow, top_has_flow=a.top_flow,
                            with_grad=bwd, search_range=a.search_range)
    cast = (lambda t: t.half()) if a.dtype == "fp16" else (lambda t: t)  # the reference's fp16 regime computes in fp16
    levels = [(cast(torch.from_numpy(d["c1"])), cast(torch.from_numpy(d["c2"])),
               None if d["flow"] is None else cast(torch.from_numpy(d["flow"])), d["scaler"]) for d in lv]
    grads = [cast(torch.from_numpy(d["g"])) for d in lv] if bwd else None
    times = []
    t_end = time.perf_counter() + budget_s
    for i in range(reps):
        t0 = time.perf_counter()
        if bwd:  # TF autodiff of the same graph (model_pwcnet.py:547-550) -> torch autograd of the port
            req = [tuple(None if t is None else t.detach().requires_grad_() for t in L[:3]) + (L[3],) for L in levels]
            outs = T.pyramid_forward(req, a.search_range)
            torch.autograd.backward(outs, grads)
        else:
            with torch.no_grad():
                T.pyramid_forward(levels, a.search_range)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() > t_end and len(times) >= 2:
            break
    return times, torch.get_num_threads()


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    if a.mode == "infer":
        return run_infer_reference(a)
    if a.mode in ("allreduce", "train"):
        print(json.dumps({"impl": "reference", "unavailable": "the reference's gradient averaging is a TF graph op on the "
                          "controller device (multi_gpus.py:68-86); there is no CPU-runnable collective to time"}), flush=True)
        return 0
    # Each step processes the arm's whole per-step batch when that keeps the run within a few minutes, else a bounded
    # sample of it (the rate is per pair either way).
    want = total_batch(a, world)
    per_pair_guess = 0.03 * (a.height * a.width) / (448 * 1024) * ((2 * a.search_range + 1) ** 2 / 81) * \
        (4.0 if a.mode == "fwdbwd" else 1.0) * (3.0 if a.dtype == "fp16" else 1.0)
    sample_batch = max(1, min(want, int(150.0 / max(1, a.warmup + a.steps) / per_pair_guess)))
    times, cores = cpu_port_time(a, sample_batch, a.warmup + a.steps, budget_s=240.0)
    timed = times[a.warmup:] if len(times) > a.warmup else times[-1:]
    ms = 1e3 * sum(timed) / len(timed)
    val = sample_batch / (ms / 1e3)
    sample = (f"{sample_batch} of the {want} pairs per step, {len(timed)} timed steps, op-for-op torch-CPU port of the "
              f"reference TF graph (TF 1.10 not installable), {cores} threads")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": len(timed),
            "warmup": min(a.warmup, max(0, len(times) - len(timed))), "ms_per_step": ms, "higher_is_better": True,
            "scaling": a.scaling, "vs_baseline": None, "dtype": "f32" if a.dtype == "fp32" else "f16",
            "data": "synthetic", "config": workload_config(a, world), "sample_pairs_per_step": sample_batch,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def library_options(_ffi):
    out = {}
    for k, dflt in OPTION_DEFAULTS.items():
        try:
            v = _ffi.get_option(k)
        except Exception:
            continue
        if dflt is None or v != dflt:
            out[k] = v
    return out


class BackwardPlan:
    """Pre-allocated buffers + one `pwc_warp_cost_volume_bwd` call per level (what autograd's backward runs)."""

    def __init__(self, dev_levels, outs, grads, r):
        import torch
        from tfoptflow_b200 import _ffi
        from tfoptflow_b200._tensors import dtype_code, ptr
        self.lib, self.calls, self.keep = _ffi.load(), [], []
        self.check = _ffi.check
        for L, out, g in zip(dev_levels, outs, grads):
            c1, c2, flow = L["c1"], L["c2"], L["flow"]
            B, H, W, C = c1.shape
            dc1, dc2 = torch.empty_like(c1), torch.empty_like(c2)
            dfl = torch.empty_like(flow) if flow is not None else None
            nb = self.lib.pwc_bwd_workspace_bytes(B, H, W, C, r, dtype_code(c1), 1 if flow is not None else 0)
            ws = torch.empty(max(nb, 256), dtype=torch.uint8, device=c1.device)
            self.keep.append((dc1, dc2, dfl, ws, out, g))
            self.calls.append((ptr(g), ptr(out), ptr(c1), ptr(c2), ptr(flow), float(L["scaler"]), ptr(dc1), ptr(dc2), ptr(dfl),
                               B, H, W, C, r, dtype_code(c1), dtype_code(flow) if flow is not None else dtype_code(c1),
                               0, 0, ptr(ws), nb))

    def run(self, i, stream):
        self.check(self.lib.pwc_warp_cost_volume_bwd(*self.calls[i], stream))


def run_ours(a):
    import torch
    import torch.distributed as dist
    from tfoptflow_b200 import _ffi, multi_gpus, pyramid, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU port)")
    rank, world = multi_gpus.init_from_env()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if a.mode == "allreduce":
        return run_allreduce(a, rank, world, local, dev)
    if a.mode == "train":
        return run_train(a, rank, world, local, dev)
    if a.mode == "infer":
        return run_infer(a, rank, world, local, dev)
    _ffi.load()
    if a.fwd_impl:
        _ffi.set_option("fwd_impl", a.fwd_impl)
    for kv in a.opt:
        k, v = kv.split("=")
        _ffi.set_option(k, int(v))
    opts = library_options(_ffi)
    if opts.get("stream_debug", 0) & 7 or opts.get("bwd_debug", 0):
        raise SystemExit("bench.py: stream_debug bits 1/2/4 and bwd_debug skip work inside the kernels (results are wrong); "
                         "refusing to emit a benchmark line")

    ndt = np_dtype(a.dtype)
    elt = 4 if a.dtype == "fp32" else 2
    r = a.search_range
    D = (2 * r + 1) ** 2
    bwd = a.mode == "fwdbwd"
    if a.scaling == "strong":
        if a.batch % world:
            raise SystemExit(f"--scaling strong: batch {a.batch} is not divisible by {world} GPUs")
        b0, b1 = multi_gpus.shard_batch(a.batch, rank, world)  # tower slicing, model_pwcnet.py:285-286
        my_batch = b1 - b0
    else:
        my_batch = a.batch
    host_levels = synth.make_pyramid(my_batch, a.height, a.width, ndt, levels=a.levels, flow=a.flow,
                                     top_has_flow=a.top_flow, with_grad=bwd, search_range=r)
    dev_levels = []
    for d in host_levels:
        dev_levels.append({"lvl": d["lvl"], "c1": torch.from_numpy(d["c1"]).to(dev), "c2": torch.from_numpy(d["c2"]).to(dev),
                           "flow": None if d["flow"] is None else torch.from_numpy(d["flow"]).to(dev),
                           "scaler": d["scaler"]})
    fwd_bytes, bwd_bytes = [], []
    for d in host_levels:
        b, h, w, c = d["c1"].shape
        hf = d["flow"] is not None
        fwd_bytes.append(b * h * w * (2 * c + (2 if hf else 0) + D) * elt)
        bwd_bytes.append(b * h * w * (2 * D + 4 * c + (4 if hf else 0)) * elt)
    dom = int(np.argmax(fwd_bytes))  # pyramid level 2 (level 1 in the operator sweep)
    # The step is two C calls: the dominant level alone (so its kernel can be bracketed by events on the launching
    # stream inside the timed region) and pwc_pyramid_fwd over all the other levels.  --per-level: one call per level.
    if a.per_level or bwd:
        plans = [pyramid.PyramidPlan([L], r) for L in dev_levels]
        dom_plan = dom
    else:
        plans = [pyramid.PyramidPlan([dev_levels[dom]], r),
                 pyramid.PyramidPlan([L for i, L in enumerate(dev_levels) if i != dom], r)]
        dom_plan = 0
    bplan = None
    if bwd:
        grads = [torch.from_numpy(d["g"]).to(dev) for d in host_levels]
        bplan = BackwardPlan(dev_levels, [p.outs[0] for p in plans], grads, r)

    def barrier():
        if world > 1:
            dist.barrier()

    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream

    def step(rec=None):
        for i, p in enumerate(plans):
            if rec is not None and (a.per_level or (i == dom_plan and not bwd)):
                rec[i][0].record(stream)
            p.run()
            if rec is not None and (a.per_level or (i == dom_plan and not bwd)):
                rec[i][1].record(stream)
        if bwd:
            for i in range(len(plans)):
                if rec is not None and (a.per_level or i == dom):
                    rec[len(plans) + i][0].record(stream)
                bplan.run(i, sp)
                if rec is not None and (a.per_level or i == dom):
                    rec[len(plans) + i][1].record(stream)

    for _ in range(max(a.warmup, 3)):
        step()
    torch.cuda.synchronize(dev)

    nrec = len(plans) * (2 if bwd else 1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    recs = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nrec)]
            for _ in range(a.steps)]
    launches0 = _ffi.launch_count()
    sampler = ClockSampler(local)
    barrier()
    torch.cuda.synchronize(dev)
    with sampler:
        ev0.record(stream)
        for s in range(a.steps):
            step(recs[s])
        ev1.record(stream)
        torch.cuda.synchronize(dev)
    barrier()
    launches = _ffi.launch_count() - launches0
    ms_step = ev0.elapsed_time(ev1) / a.steps

    def mean_ms(i):
        return sum(recs[s][i][0].elapsed_time(recs[s][i][1]) for s in range(a.steps)) / a.steps

    dom_idx = (len(plans) + dom) if bwd else dom_plan
    dom_ms = mean_ms(dom_idx)
    if a.per_level and rank == 0:
        for i in range(len(plans)):
            p = mean_ms(i)
            sys.stderr.write(f"[bench] fwd level {host_levels[i]['lvl']}: {p * 1e3:9.1f} us  "
                             f"{fwd_bytes[i] / (p * 1e-3) / 1e9:8.1f} GB/s algorithmic\n")
        for i in range(len(plans) if bwd else 0):
            p = mean_ms(len(plans) + i)
            sys.stderr.write(f"[bench] bwd level {host_levels[i]['lvl']}: {p * 1e3:9.1f} us  "
                             f"{bwd_bytes[i] / (p * 1e-3) / 1e9:8.1f} GB/s algorithmic\n")
    if world > 1:
        t = torch.tensor([ms_step, dom_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, dom_ms = float(t[0]), float(t[1])
    pairs_step = total_batch(a, world)
    value = pairs_step / (ms_step / 1e3)

    peak, peak_src = peaks()
    dom_bytes = bwd_bytes[dom] if bwd else fwd_bytes[dom]
    achieved = dom_bytes / (dom_ms / 1e3) / 1e9
    shp = host_levels[dom]["c1"].shape
    dom_flops = shp[0] * shp[1] * shp[2] * 2 * D * shp[3] * (2 if bwd else 1)
    t_hbm, t_fma = dom_bytes / (peak * 1e9), dom_flops / (FMA_PEAK_TFLOPS * 1e12)
    step_bytes = sum(fwd_bytes) + (sum(bwd_bytes) if bwd else 0)
    key = f"{'bwd_' if bwd else ''}{a.dtype}_b{my_batch}_{a.height}x{a.width}" + (f"_r{r}" if r != 4 else "")
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": roofline_traffic().get(key), "peak_source": peak_src,
            "kernel": f"fused warp+cost-volume {'backward (K1 dc1 + K2 dc2/dflow)' if bwd else 'forward'}, pyramid level "
                      f"{host_levels[dom]['lvl']} ({shp[1]}x{shp[2]}x{shp[3]})",
            "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms": dom_ms, "share_of_step": dom_ms / ms_step,
            "whole_step_frac": (step_bytes / (ms_step / 1e3) / 1e9) / peak,
            # tensor cores are ruled out for this banded correlation, so the CUDA-core FMA rate is a
            # second bound; the binding one is the larger of the two times
            "fma": {"peak_tflops": FMA_PEAK_TFLOPS, "achieved_tflops": dom_flops / (dom_ms / 1e3) / 1e12,
                    "hbm_bound_ms": t_hbm * 1e3, "fma_bound_ms": t_fma * 1e3,
                    "binding": "fma" if t_fma > t_hbm else "hbm", "frac_of_min_bound": max(t_hbm, t_fma) / (dom_ms / 1e3)}}

    # ---- spot check of the timed outputs against the oracle (2 pairs of the dominant level) ----
    check = None
    if not a.no_check and rank == 0:
        from oracle import pwc_oracle as O  # checker only
        n = min(2, my_batch)
        hd = host_levels[dom]
        tol = 1e-5 if a.dtype == "fp32" else 1e-2

        def rel(got, ref):
            ref = ref.astype(np.float64)
            return float(np.abs(got.astype(np.float64) - ref).max() / max(np.sqrt(np.mean(ref * ref)), 1e-30))
        fl = None if hd["flow"] is None else hd["flow"][:n]
        ref = O.warp_cost_volume(hd["c1"][:n], hd["c2"][:n], fl, r, hd["scaler"])
        check = {"pairs": n, "level": hd["lvl"], "tol": tol,
                 "fwd_max_err_over_rms": rel(plans[dom_plan].outs[0][:n].float().cpu().numpy(), ref)}
        if bwd and a.dtype == "fp32":
            e1, e2, e3 = O.warp_cost_volume_bwd(hd["g"][:n], hd["c1"][:n], hd["c2"][:n], fl, r, hd["scaler"])
            dc1, dc2, dfl = bplan.keep[dom][:3]
            check["dc1_max_err_over_rms"] = rel(dc1[:n].cpu().numpy(), e1)
            check["dc2_max_err_over_rms"] = rel(dc2[:n].cpu().numpy(), e2)
            if e3 is not None:
                check["dflow_max_err_over_rms"] = rel(dfl[:n].cpu().numpy(), e3)
        check["ok"] = all(v <= 2 * tol for k, v in check.items() if k.endswith("_over_rms"))
        if not check["ok"]:
            raise SystemExit(f"bench.py: timed outputs disagree with the oracle: {check}")

    # ---- end to end with HOST buffers ----
    e2e = None
    if not a.no_e2e:
        def pinned(arr):
            return torch.from_numpy(arr).pin_memory()
        n_e2e = a.e2e_steps or min(a.steps, 20)
        ceiling = None
        if not bwd:  # pwc_pyramid_fwd_host: H2D, kernels and D2H overlapped on three streams
            keep, hl, houts = [], [], []
            for d in host_levels:
                e = {"scaler": d["scaler"]}
                for k in ("c1", "c2", "flow"):
                    e[k] = None
                    if d[k] is not None:
                        keep.append(pinned(d[k]))
                        e[k] = keep[-1].numpy()
                hl.append(e)
                b, h, w, c = d["c1"].shape
                keep.append(torch.empty((b, h, w, D), dtype=keep[-1].dtype, pin_memory=True))
                houts.append(keep[-1].numpy())
            hp = pyramid.HostPyramid()
            h2d = sum(x.nbytes for e in hl for x in (e["c1"], e["c2"], e["flow"]) if x is not None)
            d2h = sum(o.nbytes for o in houts)
            for _ in range(3):
                hp.run(hl, houts, r)
            barrier()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                hp.run(hl, houts, r)  # returns after the last D2H copy has completed
            t_e2e = (time.perf_counter() - t0) / n_e2e
            path = "pwc_pyramid_fwd_host: pinned host buffers -> H2D -> kernels -> D2H, 3 streams"
            hp.close()
            # ceiling of this path: the same bytes as bare copies (uploads on one stream, downloads on another, all
            # ranks at once, no kernels) - what the host link gives N concurrent processes on this box
            s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            ups = [(L[k], torch.from_numpy(e[k])) for L, e in zip(dev_levels, hl) for k in ("c1", "c2", "flow") if e[k] is not None]
            devouts = [o for p_ in plans for o in p_.outs]  # level shapes are distinct: pair host / device outputs by shape
            dns = [(torch.from_numpy(o), next(d_ for d_ in devouts if tuple(d_.shape) == o.shape)) for o in houts]

            def bare():
                with torch.cuda.stream(s_up):
                    for dst, src in ups:
                        dst.copy_(src, non_blocking=True)
                with torch.cuda.stream(s_dn):
                    for dst, src in dns:
                        dst.copy_(src, non_blocking=True)
                s_up.synchronize()
                s_dn.synchronize()
            bare()
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                bare()
            t_bare = (time.perf_counter() - t0) / 3
            if world > 1:
                t = torch.tensor([t_bare], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                t_bare = float(t[0])
            ceiling = {"bare_copy_ms": t_bare * 1e3, "per_gpu_gbs": (h2d + d2h) / t_bare / 1e9,
                       "aggregate_gbs": world * (h2d + d2h) / t_bare / 1e9}
        else:  # training-shaped call: inputs and upstream gradients up, gradients down, through the C ABI
            hin = [{k: (None if d[k] is None else pinned(d[k])) for k in ("c1", "c2", "flow", "g")} for d in host_levels]
            hgr = [tuple(None if t is None else torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in k[:3])
                   for k in bplan.keep]
            h2d = sum(t.numel() * t.element_size() for e in hin for t in e.values() if t is not None)
            d2h = sum(t.numel() * t.element_size() for e in hgr for t in e if t is not None)

            def e2e_step():
                for L, e, g in zip(dev_levels, hin, [k[5] for k in bplan.keep]):
                    for k in ("c1", "c2", "flow"):
                        if e[k] is not None:
                            L[k].copy_(e[k], non_blocking=True)
                    g.copy_(e["g"], non_blocking=True)
                step()
                for k, hg in zip(bplan.keep, hgr):
                    for src, dst in zip(k[:3], hg):
                        if src is not None:
                            dst.copy_(src, non_blocking=True)
                torch.cuda.synchronize(dev)
            for _ in range(2):
                e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                e2e_step()
            t_e2e = (time.perf_counter() - t0) / n_e2e
            path = "pinned host inputs + upstream gradients -> H2D -> fwd + bwd kernels -> D2H of dc1/dc2/dflow, one stream"
        if world > 1:
            t = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_e2e = float(t[0])
        e2e = {"value": pairs_step / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "steps": n_e2e, "ms_per_step": t_e2e * 1e3, "path": path,
               "host_link_gbs": (h2d + d2h) / t_e2e / 1e9}
        if ceiling is not None:  # the roofline of the end-to-end path is the host link, not HBM
            e2e["host_link_ceiling"] = ceiling
            e2e["frac_of_host_link_ceiling"] = ceiling["bare_copy_ms"] / (t_e2e * 1e3)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        times, cores = cpu_port_time(a, 1, 8, budget_s=25.0)
        best = min(times[1:]) if len(times) > 1 else times[0]
        cpu = {"value": 1.0 / best, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"1 pair of the same workload ({a.height}x{a.width}, {len(a.levels)} levels, {a.dtype}, {a.mode}), "
                         f"best of {max(1, len(times) - 1)} after 1 warm-up; op-for-op torch-CPU port of the reference TF graph"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
                "dtype": "f32" if a.dtype == "fp32" else "f16", "data": "synthetic",
                "config": workload_config(a, world),
                "library": {"options": opts, "pairs_per_gpu_per_step": my_batch, "sharding": "batch-sharded, no collective",
                            "l2_policy": f"inputs larger than L2: {step_bytes / 1e6:.0f} MB of distinct tensors stream "
                                         f"through a 126 MB L2 every step" if step_bytes > 2 * 126e6 else
                                         f"{step_bytes / 1e6:.0f} MB per step: L2-resident at this per-GPU batch (not flushed)"},
                "clocks": sampler.summary(), "roofline": roof, "gpu_launches": int(launches)}
        if check is not None:
            line["oracle_check"] = check
        if e2e is not None:
            line["e2e"] = e2e
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def _infer_line(a, world, value, ms, extra):
    line = {"metric": "inference pairs/sec, pwcnet-%s-6-2 full model (image pairs per second)" % a.params, "value": value,
            "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16" if a.dtype == "fp16" else "f32",
            "data": "synthetic",
            "config": {"workload": f"BASELINE.json configs[0]: pwcnet-{a.params}-6-2 forward, {a.height}x{a.width}, batch {a.batch}"
                                   f"/GPU, random-init weights, search_range {a.search_range}, "
                                   f"{'fp16 autocast' if a.dtype == 'fp16' else 'fp32'}", "pairs_per_step": world * a.batch}}
    line.update(extra)
    return line


def run_infer(a, rank, world, local, dev):
    """BASELINE configs[0]: the whole PWC-Net forward (tfoptflow_b200/model_pwcnet.py: cuDNN convolutions + this library's
    fused warp + cost volume at every pyramid level) on synthetic image pairs with random-init weights."""
    import torch
    import torch.distributed as dist
    from tfoptflow_b200 import _ffi, model_pwcnet as M
    _ffi.load()
    mixed = a.dtype == "fp16"
    opts = M.pwcnet_options(a.params, use_mixed_precision=mixed, search_range=a.search_range)
    torch.manual_seed(1969)
    model = M.ModelPWCNet(opts).to(dev).to(memory_format=torch.channels_last).eval()
    g = torch.Generator(device=dev).manual_seed(1969 + rank)
    x = torch.rand((a.batch, 2, a.height, a.width, 3), device=dev, generator=g)
    xh = x.cpu().pin_memory()
    yh = torch.empty((a.batch, a.height, a.width, 2), dtype=torch.float32).pin_memory()

    def step(from_host=False):
        if from_host:
            x.copy_(xh, non_blocking=True)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16, enabled=mixed):
            flow_pred, _ = model(x)
        if from_host:
            yh.copy_(flow_pred.float(), non_blocking=True)
            torch.cuda.synchronize(dev)
        return flow_pred

    for _ in range(max(a.warmup, 3)):
        step()
    torch.cuda.synchronize(dev)
    launches0 = _ffi.launch_count()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with sampler:
        e0.record()
        for _ in range(a.steps):
            out = step()
        e1.record()
        torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / a.steps
    launches = _ffi.launch_count() - launches0
    n_e2e = a.e2e_steps or min(a.steps, 20)
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        step(from_host=True)
    t_e2e = (time.perf_counter() - t0) / n_e2e
    if world > 1:
        t = torch.tensor([ms, t_e2e * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, t_e2e = float(t[0]), float(t[1]) / 1e3
    if rank == 0:
        print(json.dumps(_infer_line(a, world, world * a.batch / (ms / 1e3), ms, {
            "flow_pred_finite": bool(torch.isfinite(out).all()), "clocks": sampler.summary(), "gpu_launches": int(launches),
            "e2e": {"value": world * a.batch / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(xh.numel() * 4),
                    "d2h_bytes_per_step": int(yh.numel() * 4), "steps": n_e2e, "ms_per_step": t_e2e * 1e3,
                    "path": "pinned host image pairs -> H2D -> model forward -> flow D2H"}})), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_infer_reference(a):
    """CPU arm of --mode infer: the same torch model on the host cores with the op-for-op port of the reference's
    warp + cost-volume graph (oracle/ref_graph_torch.py) as the operator - TF 1.10 itself is not installable."""
    import torch
    from oracle import ref_graph_torch as T
    from tfoptflow_b200 import model_pwcnet as M
    torch.set_num_threads(os.cpu_count() or 1)
    opts = M.pwcnet_options(a.params, search_range=a.search_range)

    def corr_fn(c1, c2, up_flow, scaler, lvl):
        warp = c2 if up_flow is None else T.dense_image_warp(c2, up_flow * scaler)
        return T.cost_volume(c1, warp, a.search_range)

    torch.manual_seed(1969)
    model = M.ModelPWCNet(opts, corr_fn=corr_fn).eval()
    x = torch.rand((a.batch, 2, a.height, a.width, 3))
    times = []
    t_end = time.perf_counter() + 200.0
    with torch.no_grad():
        for i in range(a.warmup + a.steps):
            t0 = time.perf_counter()
            model(x)
            times.append(time.perf_counter() - t0)
            if time.perf_counter() > t_end and len(times) >= 2:
                break
    timed = times[a.warmup:] if len(times) > a.warmup else times[-1:]
    ms = 1e3 * sum(timed) / len(timed)
    val = a.batch / (ms / 1e3)
    cores = torch.get_num_threads()
    line = _infer_line(a, a.gpus, val, ms, {
        "impl": "reference", "steps": len(timed),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{a.batch} pair(s) per step, {len(timed)} timed steps, torch-CPU model with the op-for-op port "
                                   f"of the reference operator graph, {cores} threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
    print(json.dumps(line), flush=True)
    return 0


def run_train(a, rank, world, local, dev):
    """BASELINE configs[3]: one training step of pwcnet-sm-6-2 per iteration - forward (cuDNN convolutions + this
    library's fused warp + cost volume at every level), pwcnet_loss, backward (this library's fused backward), the
    reference's per-rank unscale 1/128 + clip 5.0, ONE NCCL all-reduce(avg) of the flat gradient bucket, Adam
    (model_pwcnet.py:254-339).  Mixed precision = fp16 activations under autocast with fp32 master weights
    (mixed_precision.py:23-37).  Random-init weights, synthetic image pairs and flows."""
    import torch
    import torch.distributed as dist
    from tfoptflow_b200 import _ffi, model_pwcnet as M, multi_gpus
    _ffi.load()
    mixed = a.dtype == "fp16"
    opts = M.pwcnet_options(a.params, use_mixed_precision=mixed)
    torch.manual_seed(1969)
    model = M.ModelPWCNet(opts).to(dev).to(memory_format=torch.channels_last)
    optim = torch.optim.Adam(model.parameters(), lr=1e-4, eps=1e-4 if mixed else 1e-8)  # model_pwcnet.py:266-270
    params = [p for p in model.parameters()]
    g = torch.Generator(device=dev).manual_seed(1969 + rank)
    x = torch.rand((a.batch, 2, a.height, a.width, 3), device=dev, generator=g)
    y = torch.randn((a.batch, a.height, a.width, 2), device=dev, generator=g) * 4.0
    xh, yh = x.cpu().pin_memory(), y.cpu().pin_memory()
    bucket = None
    scale = float(opts["loss_scaler"]) if mixed else None

    def step(from_host=False):
        nonlocal bucket
        if from_host:
            x.copy_(xh, non_blocking=True)
            y.copy_(yh, non_blocking=True)
        optim.zero_grad(set_to_none=False)
        with torch.autocast("cuda", dtype=torch.float16, enabled=mixed):
            flow_pred, flow_pyr = model(x)
        loss = M.pwcnet_loss(y, flow_pyr, opts) + M.l2_regulariser(model, opts["gamma"])
        (loss * scale if mixed else loss).backward()
        grads = [p.grad for p in params]
        if bucket is None:
            bucket = multi_gpus.GradientBucket(grads)
        bucket.average(grads, loss_scale=scale, clip_norm=5.0 if mixed else None)
        optim.step()
        return loss

    for _ in range(max(a.warmup, 3)):
        step()
    torch.cuda.synchronize(dev)
    launches0 = _ffi.launch_count()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with sampler:
        e0.record()
        for _ in range(a.steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / a.steps
    launches = _ffi.launch_count() - launches0
    n_e2e = a.e2e_steps or min(a.steps, 10)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        lv = float(step(from_host=True))  # reads the loss back: D2H inside the timed region
    t_e2e = (time.perf_counter() - t0) / n_e2e
    if world > 1:
        t = torch.tensor([ms, t_e2e * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, t_e2e = float(t[0]), float(t[1]) / 1e3
    if rank == 0:
        print(json.dumps({
            "metric": "training samples/sec, pwcnet-%s-6-2 (image pairs per second)" % a.params, "value": world * a.batch / (ms / 1e3),
            "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16" if mixed else "f32",
            "data": "synthetic",
            "config": {"workload": f"BASELINE.json configs[3]: pwcnet-{a.params}-6-2 training step, {a.height}x{a.width}, batch "
                                   f"{a.batch}/GPU, {'mixed precision (fp16 autocast, fp32 master weights, loss scale 128, clip 5.0)' if mixed else 'fp32'}, "
                                   f"data-parallel, one NCCL all-reduce(avg) of {sum(p.numel() for p in params)} fp32 gradients, Adam",
                       "pairs_per_step": world * a.batch},
            "published_context": "reference trains sm at 52.3 -> 38.1 samples/s on Titan X + 1080 Ti (2 x 16, 256x448, fp32; "
                                 "pwcnet_train_sm-6-2-cyclic-chairsthingsmix.ipynb:299,1015) - other hardware and crop, not comparable",
            "final_loss": lv, "clocks": sampler.summary(), "gpu_launches": int(launches),
            "e2e": {"value": world * a.batch / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(xh.numel() * 4 + yh.numel() * 4),
                    "d2h_bytes_per_step": 4, "steps": n_e2e, "ms_per_step": t_e2e * 1e3,
                    "path": "pinned host batch -> H2D -> training step -> loss read back"}}), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_allreduce(a, rank, world, local, dev):
    """The training configuration's only exchange step: the gradient mean of tfoptflow/multi_gpus.py:68-86 as one NCCL
    all-reduce(avg) over a persistent flat fp32 bucket, after the per-rank unscale (1/128) and global-norm clip (5.0) of
    the reference's fp16 path (model_pwcnet.py:311-314, applied per tower BEFORE averaging)."""
    import torch
    import torch.distributed as dist
    from tfoptflow_b200 import multi_gpus
    n = {"sm": 4705064, "lg": 14079050}[a.params]  # README.md:55
    g = torch.Generator(device=dev).manual_seed(1969 + rank)
    # ~100 variables of uneven size, like the model's conv kernels and biases
    sizes, left = [], n
    while left > 0:
        s = min(left, max(16, int(n * 0.02 * (1 + (len(sizes) % 7)) / 4)))
        sizes.append(s)
        left -= s
    grads = [torch.randn(s, device=dev, generator=g) * 128.0 for s in sizes]
    bucket = multi_gpus.GradientBucket(grads)
    for _ in range(max(a.warmup, 3)):
        bucket.average(grads, loss_scale=128.0, clip_norm=5.0)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        bucket.average(grads, loss_scale=128.0, clip_norm=5.0)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / a.steps
    # the collective alone
    flat = bucket.flat
    e0.record()
    for _ in range(a.steps):
        if world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
    e1.record()
    torch.cuda.synchronize(dev)
    ms_coll = e0.elapsed_time(e1) / a.steps
    if world > 1:
        t = torch.tensor([ms, ms_coll], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_coll = float(t[0]), float(t[1])
    nbytes = n * 4
    busbw = (2 * (world - 1) / world) * nbytes / (ms_coll / 1e3) / 1e9 if world > 1 and ms_coll > 0 else 0.0
    if rank == 0:
        print(json.dumps({
            "metric": "gradient exchange steps/sec (all-reduce mean of the model's gradients)", "value": 1e3 / ms,
            "unit": "exchanges/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"multi_gpus.average_gradients: pwcnet-{a.params}-6-2 gradients, {n} fp32 = "
                                   f"{nbytes / 1e6:.1f} MB, unscale 1/128 + clip 5.0 per rank, then NCCL all-reduce(avg)"},
            "collective": {"ms": ms_coll, "bus_gbs": busbw, "nvlink_peer_gbs_measured": 770.0,
                           "frac_of_peer_bw": busbw / 770.0, "backend": "nccl" if world > 1 else "none (1 rank)"},
            "gpu_launches": 0}), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    a = parse_args()
    if a.impl == "reference":
        return run_reference(a)
    return run_ours(a)


if __name__ == "__main__":
    sys.exit(main())
