#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's configs, one JSON line on rank 0.

  --workload c2 (default; the driver's line)
      BASELINE configs[1]+[2]: 1080p synthetic clip, full resolution, every frame (I=1), all 7 complexity
      metrics + PSNR+SSIM of the same frames, on 1 B200 (N > 1: one clip-length frame range per rank of a
      virtual N-clip stream, previous rank's last frame as halo over NCCL, one all-reduce: WEAK scaling).
  --workload c1   configs[0]: the reference's own CPU case (300 x 1080p, frame_interval 10, resize 64x64),
                  identical clip in both arms (same `config.workload` string), 8-tuples printed.
  --workload c3   configs[2]: PSNR+SSIM only, 300 yuv420p 1080p pairs.
  --workload c4   configs[3]: ONE 3840x2160 clip of 3600 frames cut into contiguous frame ranges over the
                  ranks (one-frame halo, one NCCL all-reduce of the partial sums): STRONG scaling.
  --workload c5   configs[4]: 64 clips x 600 frames of 1080p, whole clips placed on ranks (clip-first plan),
                  ONE all-reduce of [64 x 7] partial sums: fixed total work.

  step     one pass of the hot path over the workload.
  value    whole-job units/s with the inputs resident in HBM when the timed region starts (device pointers
           through the C ABI); e2e = the same call with pinned HOST buffers (H2D of the planes and D2H of the
           result rows inside the timing).
  input    what the reference's flow holds per clip: the source and its encode as yuv420p planes.  PSNR/SSIM
           compares them; the complexity metrics run on the decode of the encode -- the library converts the
           encode's planes to BGR on the device exactly as cv2.VideoCapture/libswscale would,
           so ONE upload of 3 bytes per pixel and frame pair feeds both halves (round 1 uploaded 6).

`--impl reference` times the CPU arm (oracle/ref_port.py: the reference's call structure -- pool per metric,
pickled frames -- on cv2 when importable, else the C/NumPy restatement) on a bounded sample of the same
workload (config 1 runs in full), with the same `config.workload` string.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "c1": dict(metric="1080p source frames/sec, calculate_average_scene_complexity (frame_interval 10, resize 64x64)",
               unit="frames/s", h=1080, w=1920, frames=300, scaling="weak"),
    "c2": dict(metric="1080p frames/sec, all complexity metrics+PSNR/SSIM", unit="frames/s", h=1080, w=1920, frames=300,
               scaling="weak"),
    "c3": dict(metric="1080p frame pairs/sec, PSNR+SSIM", unit="pairs/s", h=1080, w=1920, frames=300, scaling="weak"),
    "c4": dict(metric="4K frames/sec, all complexity metrics+PSNR/SSIM, one clip sharded by frame range", unit="frames/s",
               h=2160, w=3840, frames=3600, scaling="strong"),
    "c5": dict(metric="1080p frames/sec, all complexity metrics+PSNR/SSIM, 64-clip farm", unit="frames/s", h=1080, w=1920,
               frames=600, scaling="strong"),
}
ALPHA = 0.8


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def workload_string(args):
    W, H, F = args.width, args.height, args.frames
    if args.workload == "c1":
        return (f"BASELINE.json configs[0]: {W}x{H} synthetic clip of {F} frames (synth_clip seed 0), frame_interval 10, "
                "resize 64x64, calculate_average_scene_complexity")
    if args.workload == "c2":
        return (f"BASELINE.json configs[1]+[2]: {W}x{H} synthetic clip of {F} frames per GPU, every frame (frame_interval 1), "
                f"resize {W}x{H}, all 7 complexity metrics + PSNR/SSIM of {F} yuv420p pairs")
    if args.workload == "c3":
        return f"BASELINE.json configs[2]: PSNR+SSIM of {F} yuv420p {W}x{H} source/encode pairs per GPU"
    if args.workload == "c4":
        return (f"BASELINE.json configs[3]: ONE {W}x{H} synthetic clip of {F} frames, every frame, all 7 complexity metrics + "
                "PSNR/SSIM, contiguous frame ranges over the GPUs with a one-frame halo")
    return (f"BASELINE.json configs[4]: {args.clips} synthetic {W}x{H} clips x {F} frames, every frame, all 7 complexity "
            "metrics + PSNR/SSIM, clip-first sharding over the GPUs")


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """SM clock / power / throttle reasons DURING the timed region (B200_PROFILING.md), every 50 ms.
    NVML in-process (nvidia_ml_py); falls back to spawning nvidia-smi.  (Spawning nvidia-smi every
    200 ms perturbed the step time by ~10 %: each invocation holds driver locks for ~150 ms.)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
        self.rows.append((float(sm), float(mx), pw, {k for k, b in bits.items() if r & b}))

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout
        f = [x.strip() for x in out.strip().split(",")]
        if len(f) >= 7:
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            self.rows.append((float(f[0]), float(f[1]), float(f[2]),
                              {n_ for n_, v in zip(names, f[3:7]) if v.lower().startswith("active")}))

    def run(self):
        while not self.stop_flag.is_set():
            try:
                self._sample_nvml() if self.nvml else self._sample_smi()
            except Exception:
                pass
            self.stop_flag.wait(0.05 if self.nvml else 1.0)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        sm = sorted(r[0] for r in self.rows)
        reasons = sorted(set().union(*[r[3] for r in self.rows]))
        out_extra = {}
        if self.nvml:
            try:
                out_extra["power_limit_w"] = self.nvml.nvmlDeviceGetEnforcedPowerLimit(self.handle) / 1000.0
            except Exception:
                pass
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.rows[0][1], "sm_mhz_min": sm[0],
                "sm_mhz_mean": round(sum(sm) / len(sm), 1), "power_w_max": max(r[2] for r in self.rows),
                "power_w_mean": round(sum(r[2] for r in self.rows) / len(self.rows), 1), **out_extra,
                "samples": len(self.rows), "source": "nvml" if self.nvml else "nvidia-smi", "reasons": reasons}


def bind_to_gpu_cpus(local):
    """One rank per GPU: run this process (and first-touch its pinned staging memory) on the CPUs NVML reports as
    local to the GPU, so the H2D uploads of N ranks do not all read one socket's DRAM over the SMP link (round-1
    finding: per-GPU H2D fell from 41 to 23 GB/s at N = 8).  Returns a description for `config`."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * wi + b for wi, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1 and 64 * wi + b < ncpu]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return {"cpus": f"{allowed[0]}-{allowed[-1]} ({len(allowed)})", "source": "nvmlDeviceGetCpuAffinity"}
    except Exception as e:  # no NVML / no affinity support: run unbound and say so
        return {"cpus": "unbound", "source": f"{type(e).__name__}: {e}"}
    return {"cpus": "unbound", "source": "empty affinity mask"}


# ----------------------------------------------------------------------------- CPU arm
def cpu_complexity_and_fr(sample_bgr, yuv_main, yuv_ref, workers, interval=1, rw=None, rh=None, with_fr=True):
    """One bounded CPU sample: all metrics (reference call structure) + PSNR/SSIM.  Returns (seconds, engine text, tuple)."""
    from oracle import ref_port as RP
    engine = "cv2" if RP.cv2 is not None else "oracle"
    h, w = sample_bgr.shape[1:3]
    t0 = time.perf_counter()
    res = RP.average_scene_complexity(sample_bgr, rw or w, rh or h, frame_interval=interval,
                                      workers=workers if engine == "cv2" else 1, engine=engine)
    if with_fr:
        cpu_psnr_ssim(yuv_main, yuv_ref, workers)
    dt = time.perf_counter() - t0
    desc = ("oracle/ref_port.py engine=cv2 (reference call structure: pool per metric, pickled frames, cv2 %s, %d workers)"
            % (RP.cv2.__version__, workers)) if engine == "cv2" else "oracle/ref_port.py engine=oracle (C + NumPy restatement, 1 thread)"
    return dt, desc, res


def cpu_psnr_ssim(yuv_main, yuv_ref, workers):
    """FFmpeg psnr/ssim restatement (oracle/c/vqa_oracle.c through ctypes, which releases the GIL) over `workers` threads."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import c_oracle as CO
    n = yuv_main[0].shape[0]

    def one(i):
        return [(CO.plane_sse(yuv_main[c][i], yuv_ref[c][i]), CO.ssim_plane(yuv_main[c][i], yuv_ref[c][i])) for c in range(3)]

    with ThreadPoolExecutor(max_workers=max(1, workers)) as ex:
        return list(ex.map(one, range(n)))


def host_sample(args, n, seed=0, first=0):
    """First n frames of the synthetic workload on the HOST (torch CPU generator): source / encode planes and the BGR
    frames cv2.VideoCapture would decode from the encode (oracle restatement of libswscale's conversion)."""
    import rtvqa_b200  # noqa: F401
    from rtvqa_b200.synth_device import DeviceClipSynth
    from oracle import np_oracle as NO
    syn = DeviceClipSynth(args.height, args.width, seed, "cpu")
    ref, enc = syn.pairs(first, n)
    ref = [p.numpy() for p in ref]
    enc = [p.numpy() for p in enc]
    bgr = np.stack([NO.yuv420_to_bgr(enc[0][i], enc[1][i], enc[2][i]) for i in range(n)])
    return bgr, enc, ref


def run_reference(args):
    if env_int("RANK", 0) != 0:
        return 0
    from oracle import c_oracle
    c_oracle.build()
    import rtvqa_b200
    cores = os.cpu_count() or 1
    wl = WORKLOADS[args.workload]
    extra = {}
    if args.workload == "c1":
        # config 1 in FULL, at the reference's default worker count and at all cores
        clip = rtvqa_b200.synth.synth_clip(args.frames, args.height, args.width, seed=0)
        runs = {}
        for W in sorted({max(1, cores // 2), cores}):
            ts = []
            for it in range(args.warmup + args.steps):
                dt, desc, res = cpu_complexity_and_fr(clip, None, None, W, interval=10, rw=64, rh=64, with_fr=False)
                if it >= args.warmup:
                    ts.append(dt)
            runs[W] = (float(np.mean(ts)), desc, res)
        best = min(runs, key=lambda k: runs[k][0])
        per, desc, res = runs[best]
        units, sample = args.frames, f"full config: {args.frames} frames ({args.frames // 10 - 1} analysed pairs); {desc}"
        extra = {"workers_runs": {str(k): {"s_per_clip": v[0], "frames_per_s": args.frames / v[0]} for k, v in runs.items()},
                 "result": {"scene_complexity": [float(x) for x in res]}}
        used = best
    elif args.workload == "c3":
        n = max(8, min(args.ref_frames * 4, args.frames))
        _, enc, ref = host_sample(args, n, seed=1)
        ts = []
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            cpu_psnr_ssim(enc, ref, cores)
            if it >= args.warmup:
                ts.append(time.perf_counter() - t0)
        per, units, used = float(np.mean(ts)), n, cores
        sample = f"{n} pairs/step; oracle/c/vqa_oracle.c (FFmpeg psnr/ssim restatement) over {cores} threads"
    else:
        n = max(9, min(args.ref_frames if args.workload != "c4" else max(9, args.ref_frames // 2), args.frames))
        bgr, enc, ref = host_sample(args, n, seed={"c2": 0, "c4": 2, "c5": 100}[args.workload])
        ts = []
        for it in range(args.warmup + args.steps):
            dt, desc, _ = cpu_complexity_and_fr(bgr, enc, ref, cores)
            if it >= args.warmup:
                ts.append(dt)
        per, units, used = float(np.mean(ts)), n - 1, cores
        sample = f"{n} frames/step ({n - 1} analysed); {desc}"
    value = units / per
    import platform
    cpu_model = ""
    try:
        with open("/proc/cpuinfo") as f:
            cpu_model = next((ln.split(":", 1)[1].strip() for ln in f if ln.startswith("model name")), "")
    except Exception:
        pass
    try:
        import cv2
        cvt = cv2.getNumThreads()
    except Exception:
        cvt = None
    line = {
        "impl": "reference", "metric": wl["metric"], "value": value, "unit": wl["unit"], "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": wl["scaling"],
        "vs_baseline": None, "dtype": "u8/f32", "data": "synthetic",
        "config": {"workload": workload_string(args)},
        "cpu_baseline": {"value": value, "unit": wl["unit"], "cores": used, "kind": "port", "sample": sample,
                         "cpu_model": cpu_model or platform.processor(), "cv2_threads": cvt, "host_cores": cores},
        "e2e": {"value": value, "unit": wl["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, **extra,
    }
    _emit(line)
    return 0


# ----------------------------------------------------------------------------- GPU arm helpers
class Bench:
    """Process-level state of the GPU arm: device, context, process group."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world, self.rank, self.local = env_int("WORLD_SIZE", 1), env_int("RANK", 0), env_int("LOCAL_RANK", 0)
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback (use --impl reference)")
        self.binding = bind_to_gpu_cpus(self.local) if self.world > 1 else {"cpus": "all (single rank)", "source": "-"}
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            dist.init_process_group("nccl", device_id=self.dev)
        from rtvqa_b200 import _native as N
        from rtvqa_b200 import sharding as SH
        self.N, self.SH = N, SH
        self.ctx = N.get_context(self.local)
        self.ctx.use_torch_stream()
        if self.world > 1:
            SH.init_context_comm(self.ctx)            # the library's own NCCL communicator (vqa_clip_reduce)

    def pin(self, t):
        return t.cpu().pin_memory()

    def timed(self, fn, steps, warmup):
        torch, dist = self.torch, self.dist
        for _ in range(warmup):
            fn()
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = self.ctx.kernel_launches()
        t0 = time.perf_counter()
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if self.world > 1:
            dist.barrier()
        ms = max(e0.elapsed_time(e1), 0.0)
        # the ABI calls end with a D2H + stream sync, so device time == wall time to within the launch
        # overhead; take the larger and the max over ranks
        sec = max(ms / 1e3, wall)
        t = torch.tensor([sec], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), self.ctx.kernel_launches() - l0, out

    def roofline(self, profile_fn, units_per_profile=None):
        """Per-kernel CUDA-event pass (vqa_kernel_profile) over `profile_fn` -> the roofline object of the kernel with
        the largest share, measured live on the context's stream; peak from MEASURED_PEAKS.json."""
        ctx = self.ctx
        ctx.kernel_profile(True)
        profile_fn()
        rep = ctx.kernel_report()
        ctx.kernel_profile(False)
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        tot_ms = sum(v["ms"] for v in rep.values()) or 1.0
        tname, tv = max(rep.items(), key=lambda kv: kv[1]["ms"])
        achieved = tv["bytes"] / (tv["ms"] * 1e-3) / 1e9 if tv["ms"] > 0 else 0.0
        roof = {"bound": "hbm", "kernel": tname, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": None, "peak_source": peak_src,
                "launches_per_profile_pass": tv["launches"], "avg_launch_ms": tv["ms"] / max(tv["launches"], 1),
                "share_of_step": tv["ms"] / tot_ms,
                "algorithmic_bytes_per_launch": tv["bytes"] / max(tv["launches"], 1),
                "kernel_sum_ms": round(tot_ms, 3),
                "kernels": {k: {"ms": round(v["ms"], 4), "launches": v["launches"],
                                "GBps": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 and v["bytes"] else None,
                                "frac_of_hbm_peak": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9 / hbm_peak, 3) if v["ms"] > 0 and v["bytes"] else None,
                                "TFLOPs": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2) if v["flops"] else None}
                            for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"])}}
        traffic_file = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(traffic_file):
            try:
                with open(traffic_file) as f:
                    t = json.load(f).get(tname)
                if t and t.get("ratio"):
                    # DRAM bytes per launch = (dram bytes / algorithmic bytes of the ncu-captured launch) x this run's
                    # algorithmic bytes per launch (launches differ in size across pyramid levels)
                    roof["traffic"] = t["ratio"] * roof["algorithmic_bytes_per_launch"]
                    roof["traffic_source"] = {"report": t["report"], "captured_dram_bytes": t["dram_bytes"],
                                              "captured_algorithmic_bytes": t["algorithmic_bytes"], "ratio": t["ratio"]}
            except Exception:
                pass
        return roof

    def finish(self):
        if self.world > 1:
            self.dist.barrier()
            try:
                self.ctx.comm_destroy()
            except Exception:
                pass
            self.dist.destroy_process_group()


def sha_ints(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a, dtype=np.int64).tobytes())
    return h.hexdigest()[:16]


def gather_int_rows(B, rows):
    """Per-frame integer outputs of all ranks in frame order -> (edge total, orb total, digest); must be identical for
    every rank count."""
    edge = np.asarray(rows["edge_count"], dtype=np.int64)
    orb = np.asarray(rows["orb_count"], dtype=np.int64)
    if B.world > 1:
        box = [None] * B.world
        B.dist.all_gather_object(box, (edge, orb))
        edge = np.concatenate([b[0] for b in box])
        orb = np.concatenate([b[1] for b in box])
    return edge, orb


def base_line(B, args, value, sec, steps, launches, clocks, roof, cpu, e2e, result, config_extra):
    wl = WORKLOADS[args.workload]
    return {
        "metric": wl["metric"], "value": value, "unit": wl["unit"], "n_gpus": B.world, "steps": steps, "warmup": args.warmup,
        "ms_per_step": sec / steps * 1e3, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
        "dtype": "u8/f32 (integer pixel paths; fp32 flow/DCT, fp64 reductions)", "data": "synthetic",
        "config": {"workload": workload_string(args), **config_extra, "cpu_binding": B.binding},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "result": result,
    }


# ----------------------------------------------------------------------------- c2 (default) and c4
def run_frames(args):
    """c2: one clip-length range per rank (weak).  c4: ONE clip cut into frame ranges (strong).  Same step function:
    halo -> vqa_analyze_clip_yuv420 on the rank's range -> weighted partial sums -> vqa_clip_reduce -> finalize."""
    B = Bench(args)
    torch, N, SH, ctx = B.torch, B.N, B.SH, B.ctx
    import rtvqa_b200
    from rtvqa_b200.synth_device import DeviceClipSynth
    H, W = args.height, args.width
    strong = args.workload == "c4"
    if strong:
        k_total = args.frames
        a0, b0 = SH.shard_range(k_total, B.rank, B.world)
        syn = DeviceClipSynth(H, W, 2, B.dev)
    else:
        k_total = args.frames * B.world
        a0, b0 = B.rank * args.frames, (B.rank + 1) * args.frames
        syn = DeviceClipSynth(H, W, B.rank, B.dev)      # every rank its own clip: a virtual stream of N clips
    m = b0 - a0
    # ---- workload (outside every timed region)
    t_gen = time.perf_counter()
    ref_dev, enc_dev = syn.pairs(a0 if strong else 0, m)
    halo_dev = None
    if strong and a0 > 0:
        halo_dev = [p.contiguous() for p in syn.pair(a0 - 1)[1]]
    pack = torch.empty(H * W * 3 // 2, dtype=torch.uint8, device=B.dev)
    last = torch.cat([enc_dev[c][-1].reshape(-1) for c in range(3)]).contiguous()
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    ts = rtvqa_b200.synth.synth_timestamps(k_total, 60.0 if strong else 30.0)
    orb_size = tuple(int(v) for v in args.orb_size.lower().split("x")) if args.orb_size else None

    def halo_planes():
        if strong or B.world == 1:
            return halo_dev
        # weak-scaling stream: previous rank's last frame over NVLink (vqa_comm_halo_exchange, grouped send/recv)
        ctx.halo_exchange(last if B.rank + 1 < B.world else None, pack if B.rank > 0 else None)
        if B.rank == 0:
            return None
        hw = H * W
        return [pack[:hw].view(H, W), pack[hw:hw + hw // 4].view(H // 2, W // 2), pack[hw + hw // 4:].view(H // 2, W // 2)]

    def close(rows):
        partials = SH.local_partials(rows, a0, k_total, ALPHA, ctx.ewm_partial)
        lo = max(1 - a0, 0)
        ints = np.array([int(rows["edge_count"][lo:].sum()), int(rows["orb_count"][lo:].sum()), len(rows)], dtype=np.int64)
        partials, ints = SH.reduce_partials(partials, ints, ctx=ctx)
        fps = ctx.framerate_series(ts)
        return SH.finalize(partials, k_total, ctx.ewm_partial(fps, 0, len(fps), ALPHA)), ints

    def step_device():
        rows, fr = ctx.analyze_clip_yuv420(enc_dev, ref_dev, W, H, halo_planes=halo_planes(), orb_size=orb_size)
        res, ints = close(rows)
        return res, ints, rows, fr

    # e2e: the same call on pinned HOST planes.  c4 keeps the host copy bounded (a sub-range of the rank's frames)
    e_n = m if not strong else min(m, args.e2e_frames)
    enc_host = [B.pin(p[:e_n]).numpy() for p in enc_dev]
    ref_host = [B.pin(p[:e_n]).numpy() for p in ref_dev]
    halo_host = [p.cpu().numpy() for p in halo_dev] if halo_dev is not None else None

    def step_e2e():
        hp = halo_host
        if not strong and B.world > 1:
            hd = halo_planes()
            hp = [p.cpu().numpy() for p in hd] if hd is not None else None
        rows, fr = ctx.analyze_clip_yuv420(enc_host, ref_host, W, H, halo_planes=hp, orb_size=orb_size)
        vals = close(rows) if not strong else None     # strong mode: bounded sub-range, no collective in the e2e leg
        return rows, fr, vals

    sampler = ClockSampler(B.local)
    sampler.start()
    sec_dev, launches, out = B.timed(step_device, args.steps, args.warmup)
    clocks = sampler.summary()
    e2e_steps = max(1, min(args.steps, 5))
    sec_e2e, _, out_e2e = B.timed(step_e2e, e2e_steps, 2)
    total_analysed = k_total - 1
    value = total_analysed * args.steps / sec_dev
    e2e_units = e_n * B.world - 1 if strong else total_analysed
    e2e_value = e2e_units * e2e_steps / sec_e2e

    # ---- checks inside the run: (1) the e2e rows equal the device-resident rows bit for bit on the frames both saw;
    # (2) the all-gathered per-frame table smoothed on rank 0 agrees with the reduce path to <= 1e-12
    res, ints, rows, fr = out
    same = all(np.array_equal(out_e2e[0][f][:e_n], rows[f][:e_n]) for f in ("edge_count", "orb_count", "gray_sq_sum")) and \
        np.allclose(out_e2e[1]["psnr_avg"], fr["psnr_avg"][:e_n], rtol=0, atol=0)
    table = SH.gather_rows(rows, device=B.dev)
    from rtvqa_b200 import complexity_metrics as cm
    alt = SH.means_from_table(table, ALPHA, cm.smooth_data)
    rel = max(abs(a - b) / max(abs(b), 1e-300) for a, b in zip(alt, [float(v) for v in res[:7]]))
    edge_all, orb_all = gather_int_rows(B, rows)

    nprof = min(m, 96 if strong else m)
    roof = B.roofline(lambda: ctx.analyze_clip_yuv420([p[:nprof] for p in enc_dev], [p[:nprof] for p in ref_dev], W, H,
                                                      orb_size=orb_size))
    roof["end_to_end_input_GBps"] = value * 3 * H * W / 1e9 / B.world
    roof["profile_pass"] = f"{nprof} frames of this rank's range, side stream off (kernels serialised for per-kernel events)"
    if B.rank == 0:
        cpu = None
        if B.world == 1 and not args.no_cpu_baseline:
            from oracle import c_oracle, np_oracle as NO
            c_oracle.build()
            n = max(9, min(args.ref_frames if not strong else max(9, args.ref_frames // 2), m))
            cores = os.cpu_count() or 1
            enc_s = [p[:n].cpu().numpy() for p in enc_dev]
            ref_s = [p[:n].cpu().numpy() for p in ref_dev]
            bgr_s = np.stack([NO.yuv420_to_bgr(enc_s[0][i], enc_s[1][i], enc_s[2][i]) for i in range(n)])
            dt, desc, _ = cpu_complexity_and_fr(bgr_s, enc_s, ref_s, cores)
            cpu = {"value": (n - 1) / dt, "unit": "frames/s", "cores": cores, "kind": "port",
                   "sample": f"first {n} frames of the bench clip ({n - 1} analysed), {dt:.1f} s; {desc}"}
        h2d = sum(p.nbytes for p in enc_host) + sum(p.nbytes for p in ref_host)
        d2h = out_e2e[0].nbytes + out_e2e[1].nbytes
        cfgx = {"frames_per_gpu": m, "analysed_frames_per_step": total_analysed, "clip_frames": k_total,
                "input": "source + encode as yuv420p planes (3 B/px per frame pair); BGR derived on the device (libswscale-exact)",
                "l2": f"inputs ({sum(p.numel() for p in enc_dev + ref_dev) / 1e6:.0f} MB per rank and step) larger than the 126 MB L2; no flush",
                "parallelism": (f"ONE clip, contiguous frame ranges x{B.world}, one-frame halo, one vqa_clip_reduce (ncclAllReduce)"
                                if strong else f"frame-range x{B.world} of a virtual stream, one-frame halo over vqa_comm_halo_exchange, "
                                "one vqa_clip_reduce (ncclAllReduce)"),
                "workload_generation_s": round(t_gen, 1)}
        if orb_size:
            cfgx["orb_size"] = "%dx%d (extension: full ORB pipeline instead of the reference's 64x64)" % orb_size
        e2e = {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "steps": e2e_steps, "api": "vqa_analyze_clip_yuv420 on pinned host planes (one upload feeds both halves)",
               "frames_per_gpu": e_n, "rows_equal_device_resident": bool(same)}
        result = {"scene_complexity": [float(v) for v in res], "psnr_avg_first": float(fr["psnr_avg"][0]),
                  "ssim_all_first": float(fr["ssim_all"][0]),
                  "cross_n_check": {"edge_total": int(edge_all[1:].sum()), "orb_total": int(orb_all[1:].sum()),
                                    "frames": int(len(edge_all)), "int_rows_sha": sha_ints(edge_all, orb_all),
                                    "reduced_ints": [int(v) for v in ints],
                                    "gather_vs_reduce_max_rel": rel}}
        assert rel <= 1e-12, f"gathered-table means differ from the reduce path: {rel}"
        assert int(ints[0]) == int(edge_all[1:].sum()) and int(ints[1]) == int(orb_all[1:].sum()), "integer totals differ"
        _emit(base_line(B, args, value, sec_dev, args.steps, launches, clocks, roof, cpu, e2e, result, cfgx))
    B.finish()
    return 0


# ----------------------------------------------------------------------------- c3: PSNR + SSIM only
def run_c3(args):
    B = Bench(args)
    torch, ctx = B.torch, B.ctx
    from rtvqa_b200.synth_device import DeviceClipSynth
    H, W, F = args.height, args.width, args.frames
    syn = DeviceClipSynth(H, W, 1 + B.rank, B.dev)
    ref_dev, enc_dev = syn.pairs(0, F)
    enc_host = [B.pin(p).numpy() for p in enc_dev]
    ref_host = [B.pin(p).numpy() for p in ref_dev]
    torch.cuda.synchronize()
    sampler = ClockSampler(B.local)
    sampler.start()
    sec_dev, launches, fr = B.timed(lambda: ctx.psnr_ssim(enc_dev, ref_dev), args.steps, args.warmup)
    clocks = sampler.summary()
    e2e_steps = max(1, min(args.steps, 5))
    sec_e2e, _, fr_h = B.timed(lambda: ctx.psnr_ssim(enc_host, ref_host), e2e_steps, 2)
    value = F * B.world * args.steps / sec_dev
    roof = B.roofline(lambda: ctx.psnr_ssim(enc_dev, ref_dev))
    roof["end_to_end_input_GBps"] = value * 3 * H * W / 1e9 / B.world
    if B.rank == 0:
        cpu = None
        if B.world == 1 and not args.no_cpu_baseline:
            from oracle import c_oracle
            c_oracle.build()
            n, cores = min(F, 4 * args.ref_frames), os.cpu_count() or 1
            t0 = time.perf_counter()
            got = cpu_psnr_ssim([p[:n] for p in enc_host], [p[:n] for p in ref_host], cores)
            dt = time.perf_counter() - t0
            # the timed CPU rows double as a parity check of the bench's own output
            for i in (0, n - 1):
                assert [g[0] for g in got[i]] == [int(v) for v in fr["sse"][i]], "SSE differs from the CPU oracle"
            cpu = {"value": n / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
                   "sample": f"first {n} pairs, {dt:.1f} s; oracle/c/vqa_oracle.c (FFmpeg psnr/ssim restatement) over {cores} threads"}
        e2e = {"value": F * B.world * e2e_steps / sec_e2e, "unit": "pairs/s",
               "h2d_bytes_per_step": int(sum(p.nbytes for p in enc_host + ref_host)), "d2h_bytes_per_step": int(fr_h.nbytes),
               "steps": e2e_steps, "api": "vqa_psnr_ssim_planar on pinned host planes",
               "rows_equal_device_resident": bool(np.array_equal(fr_h["sse"], fr["sse"]) and np.array_equal(fr_h["ssim_all"], fr["ssim_all"]))}
        cfgx = {"pairs_per_gpu": F, "l2": f"inputs ({3 * H * W * F / 1e6:.0f} MB per step) larger than the 126 MB L2; no flush",
                "parallelism": f"independent pair ranges x{B.world}, no collective"}
        result = {"psnr_avg_first": float(fr["psnr_avg"][0]), "ssim_all_first": float(fr["ssim_all"][0]),
                  "psnr_avg_mean": float(np.mean(fr["psnr_avg"])), "ssim_all_mean": float(np.mean(fr["ssim_all"]))}
        _emit(base_line(B, args, value, sec_dev, args.steps, launches, clocks, roof, cpu, e2e, result, cfgx))
    B.finish()
    return 0


# ----------------------------------------------------------------------------- c5: clip farm
def run_c5(args):
    B = Bench(args)
    torch, N, SH, ctx = B.torch, B.N, B.SH, B.ctx
    import rtvqa_b200
    from rtvqa_b200.synth_device import DeviceClipSynth
    H, W, F, C = args.height, args.width, args.frames, args.clips
    clip_frames = [F] * C
    plan = SH.plan_clip_shards(clip_frames, B.world)[B.rank]
    need = sum(b - a for _, a, b in plan) * 3 * H * W
    free = torch.cuda.mem_get_info(B.dev)[0]
    if need > 0.8 * free:
        if B.rank == 0:
            _emit({"metric": WORKLOADS["c5"]["metric"], "n_gpus": B.world, "config": {"workload": workload_string(args)},
                   "unavailable": f"{need / 1e9:.0f} GB of resident planes per rank do not fit {free / 1e9:.0f} GB of free HBM: "
                                  "use more GPUs or --clips"})
        B.finish()
        return 0
    t_gen = time.perf_counter()
    shards = []
    for clip, a, b in plan:
        syn = DeviceClipSynth(H, W, 100 + clip, B.dev)
        ref, enc = syn.pairs(a, b - a)
        halo = [p.contiguous() for p in syn.pair(a - 1)[1]] if a > 0 else None
        shards.append((clip, a, b, ref, enc, halo))
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    ts = [rtvqa_b200.synth.synth_timestamps(F, 30.0)] * C

    def step(host=None):
        rows_by = {}

        def rows_of(clip, a, b):
            sh = next(s for s in shards if s[0] == clip and s[1] == a)
            enc, ref = (host[(clip, a)] if host else (sh[4], sh[3]))
            halo = sh[5] if not host or sh[5] is None else [p.cpu().numpy() for p in sh[5]]
            rows, fr = ctx.analyze_clip_yuv420(enc, ref, W, H, halo_planes=halo)
            rows_by[(clip, a)] = (rows, fr)
            return rows

        partials, ints = SH.multi_clip_partials(plan, rows_of, clip_frames, ALPHA, ctx.ewm_partial)
        partials, ints = SH.reduce_partials(partials, ints, ctx=ctx)
        fps = ctx.framerate_series(ts[0])
        fmean = ctx.ewm_partial(fps, 0, len(fps), ALPHA)
        return [SH.finalize(partials[c], clip_frames[c], fmean) for c in range(C)], ints, rows_by

    sampler = ClockSampler(B.local)
    sampler.start()
    sec_dev, launches, out = B.timed(step, args.steps, args.warmup)
    clocks = sampler.summary()
    # e2e: pinned host planes of (at most --e2e-clips of) this rank's shards
    e_sh = shards[:max(1, args.e2e_clips)]
    host = {(s[0], s[1]): ([B.pin(p).numpy() for p in s[4]], [B.pin(p).numpy() for p in s[3]]) for s in e_sh}
    e_plan = [(s[0], s[1], s[2]) for s in e_sh]

    def step_e2e():
        got = {}
        for clip, a, b in e_plan:
            enc, ref = host[(clip, a)]
            sh = next(s for s in e_sh if s[0] == clip and s[1] == a)
            got[(clip, a)] = ctx.analyze_clip_yuv420(enc, ref, W, H, halo_planes=[p.cpu().numpy() for p in sh[5]] if sh[5] is not None else None)
        return got

    e2e_steps = max(1, min(args.steps, 3))
    sec_e2e, _, got_e = B.timed(step_e2e, e2e_steps, 1)
    total = sum(k - 1 for k in clip_frames)
    value = total * args.steps / sec_dev
    e_frames = sum(b - a for _, a, b in e_plan)
    e2e_value = (e_frames - len(e_plan)) * B.world * e2e_steps / sec_e2e
    res, ints, rows_by = out
    k0 = (e_plan[0][0], e_plan[0][1])
    same = np.array_equal(got_e[k0][0]["edge_count"], rows_by[k0][0]["edge_count"]) and \
        np.array_equal(got_e[k0][1]["sse"], rows_by[k0][1]["sse"])
    roof = B.roofline(lambda: ctx.analyze_clip_yuv420([p[:96] for p in shards[0][4]], [p[:96] for p in shards[0][3]], W, H))
    roof["end_to_end_input_GBps"] = value * 3 * H * W / 1e9 / B.world
    if B.rank == 0:
        h2d = sum(sum(p.nbytes for p in e) + sum(p.nbytes for p in r) for e, r in host.values())
        e2e = {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(sum(v[0].nbytes + v[1].nbytes for v in got_e.values())), "steps": e2e_steps,
               "api": "vqa_analyze_clip_yuv420 on pinned host planes", "clips_per_gpu": len(e_plan),
               "rows_equal_device_resident": bool(same)}
        cfgx = {"clips": C, "frames_per_clip": F, "analysed_frames_per_step": total, "shards_of_rank0": len(plan),
                "input": "source + encode as yuv420p planes; BGR derived on the device (libswscale-exact)",
                "l2": f"inputs ({need / 1e6:.0f} MB per rank and step) larger than the 126 MB L2; no flush",
                "parallelism": f"clip-first plan over {B.world} ranks (plan_clip_shards), ONE vqa_clip_reduce of [{C} x 7] doubles + [{C} x 3] integers",
                "workload_generation_s": round(t_gen, 1)}
        flat = np.array([[float(v) for v in r] for r in res])
        result = {"clip0_scene_complexity": [float(v) for v in res[0]],
                  "cross_n_check": {"edge_total": int(ints[:, 0].sum()), "orb_total": int(ints[:, 1].sum()),
                                    "frames": int(ints[:, 2].sum()), "ints_sha": sha_ints(ints),
                                    "float_means_of_clip_means": [float(v) for v in flat.mean(axis=0)]}}
        _emit(base_line(B, args, value, sec_dev, args.steps, launches, clocks, roof, None, e2e, result, cfgx))
    B.finish()
    return 0


# ----------------------------------------------------------------------------- c1: the reference's own config
def run_c1(args):
    B = Bench(args)
    torch, N, ctx = B.torch, B.N, B.ctx
    import rtvqa_b200
    from rtvqa_b200 import complexity_metrics as cm
    H, W, F, I = args.height, args.width, args.frames, 10
    clip = rtvqa_b200.synth.synth_clip(F, H, W, seed=0)              # the clip of tests/golden (config 1 of the reference)
    sampled_host = torch.from_numpy(np.ascontiguousarray(clip[I - 1::I])).pin_memory()
    sampled_dev = sampled_host.to(B.dev)
    ts = rtvqa_b200.synth.synth_timestamps(F, 30.0)[::I]
    torch.cuda.synchronize()

    def close(rows):
        g = lambda name, first: cm._smoothed_mean(rows[name][first:], ALPHA, empty=0.0 if name == "temporal_dct" else float("nan"))
        fps = ctx.framerate_series(ts)
        return (g("motion", 1), g("dct_energy", 1), g("hist_entropy", 1), g("edge_count", 1), g("orb_count", 1),
                g("color_entropy", 1), g("temporal_dct", 2), cm._smoothed_mean(fps, ALPHA))

    sampler = ClockSampler(B.local)
    sampler.start()
    sec_dev, launches, res = B.timed(lambda: close(ctx.complexity_frames(sampled_dev, 64, 64)), args.steps, args.warmup)
    clocks = sampler.summary()
    host_np = sampled_host.numpy()
    e2e_steps = max(1, min(args.steps, 5))
    sec_e2e, _, res_h = B.timed(lambda: close(ctx.complexity_frames(host_np, 64, 64)), e2e_steps, 2)
    value = F * B.world * args.steps / sec_dev
    roof = B.roofline(lambda: ctx.complexity_frames(sampled_dev, 64, 64))
    if B.rank == 0:
        cpu = None
        if B.world == 1 and not args.no_cpu_baseline:
            from oracle import c_oracle
            c_oracle.build()
            cores = os.cpu_count() or 1
            dt, desc, ref_res = cpu_complexity_and_fr(clip, None, None, cores, interval=I, rw=64, rh=64, with_fr=False)
            cpu = {"value": F / dt, "unit": "frames/s", "cores": cores, "kind": "port",
                   "sample": f"full config: {F} frames ({F // I - 1} analysed pairs), {dt:.1f} s; {desc}",
                   "result": [float(v) for v in ref_res],
                   "max_rel_vs_gpu": max(abs(float(a) - float(b)) / max(abs(float(b)), 1e-300) for a, b in zip(res, ref_res))}
        e2e = {"value": F * B.world * e2e_steps / sec_e2e, "unit": "frames/s", "h2d_bytes_per_step": int(host_np.nbytes),
               "d2h_bytes_per_step": int(len(host_np) * N.FRAME_DTYPE.itemsize), "steps": e2e_steps,
               "api": "vqa_complexity_frames on the pinned sampled frames (the reference decodes and samples on the CPU)",
               "tuple_equal_device_resident": bool(all(float(a) == float(b) for a, b in zip(res, res_h)))}
        cfgx = {"sampled_frames": int(len(host_np)), "l2": f"inputs ({host_np.nbytes / 1e6:.0f} MB per step) larger than the 126 MB L2; no flush",
                "parallelism": f"replicas x{B.world}"}
        _emit(base_line(B, args, value, sec_dev, args.steps, launches, clocks, roof, cpu, e2e,
                        {"scene_complexity": [float(v) for v in res]}, cfgx))
    B.finish()
    return 0


_JSON_OUT = None


def _claim_stdout():
    """stdout must carry exactly one JSON line, but libraries write to fd 1 behind Python's back (NCCL
    prints its version banner with plain printf at NCCL_DEBUG=VERSION/WARN).  Keep a private handle on
    the real stdout for the JSON line and point fd 1 at stderr for everything else."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=None, help="c2/c3: frames per GPU; c4: frames of THE clip; c5: frames per clip; c1: source frames")
    ap.add_argument("--clips", type=int, default=64, help="c5: number of clips")
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--ref-frames", type=int, default=33, help="frames per CPU sample (bounded)")
    ap.add_argument("--e2e-frames", type=int, default=256, help="c4: frames per rank kept in pinned host memory for the e2e leg")
    ap.add_argument("--e2e-clips", type=int, default=2, help="c5: clips per rank kept in pinned host memory for the e2e leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--orb-size", default="", help="WxH: full ORB pipeline on gray(resize(frame, WxH)) (default: reference 64x64)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    args.frames = args.frames or wl["frames"]
    args.height = args.height or wl["h"]
    args.width = args.width or wl["w"]
    if args.steps is None:
        args.steps = {"c1": 5, "c2": 5, "c3": 10, "c4": 2, "c5": 2}[args.workload]
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3
    return {"c1": run_c1, "c2": run_frames, "c3": run_c3, "c4": run_frames, "c5": run_c5}[args.workload](args)


if __name__ == "__main__":
    sys.exit(main())
