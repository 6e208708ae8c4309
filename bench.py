#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config, one JSON line on rank 0.

  metric   1080p frames/sec, all 7 complexity metrics + PSNR/SSIM  (BASELINE.json)
  workload configs[1]+[2]: 1080p30 synthetic clip, full resolution, every frame (I=1), all 7
           complexity metrics, plus PSNR+SSIM of the same number of yuv420p pairs, on 1 B200.
  step     one pass of the hot path over one clip (default 300 frames = 299 analysed frames).
  value    analysed frames/s with the clip already resident in HBM (device pointers through the
           C ABI); e2e = the same through the reference-shaped public API with pinned HOST
           buffers (H2D of the clip + yuv planes and D2H of the result rows inside the timing).
  N > 1    one process per GPU (torchrun); every rank owns one clip-length frame range of a
           virtual N-clip stream: the previous rank's last frame arrives as a halo over NCCL
           (send/recv), ranks run with no data-path collective, and one all-reduce of the
           per-rank weighted partial sums closes the step (weak scaling).

`--impl reference` times the CPU arm instead (the oracle port with the reference's call
structure: oracle/ref_port.py, engine cv2 when OpenCV is importable on the host, else the C/NumPy
restatement) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "1080p frames/sec, all complexity metrics+PSNR/SSIM"
UNIT = "frames/s"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ----------------------------------------------------------------------------- workload
def make_clip_host(n, h, w, seed):
    import rtvqa_b200
    return rtvqa_b200.synth.synth_clip(n, h, w, seed=seed)


def make_yuv_pairs_device(clip_dev, seed):
    """Distorted twin for the full-reference half, derived on the device (workload generation,
    outside every timed region): BT.601 BGR->yuv420p, distortion = 3x3 blur + {-2..2} noise."""
    import torch
    x = clip_dev.to(torch.int32)
    b, g, r = x[..., 0], x[..., 1], x[..., 2]
    y = ((66 * r + 129 * g + 25 * b + 128) >> 8) + 16
    u = ((-38 * r - 74 * g + 112 * b + 128) >> 8) + 128
    v = ((112 * r - 94 * g - 18 * b + 128) >> 8) + 128

    def sub(p):
        return (p[:, 0::2, 0::2] + p[:, 0::2, 1::2] + p[:, 1::2, 0::2] + p[:, 1::2, 1::2] + 2) >> 2

    ref = [y.clamp(0, 255).to(torch.uint8), sub(u).clamp(0, 255).to(torch.uint8), sub(v).clamp(0, 255).to(torch.uint8)]
    gen = torch.Generator(device=clip_dev.device)
    gen.manual_seed(1234 + seed)
    k = torch.tensor([[1., 2., 1.], [2., 4., 2.], [1., 2., 1.]], device=clip_dev.device).view(1, 1, 3, 3) / 16.0
    dist = []
    for p in ref:
        q = torch.nn.functional.pad(p.float().unsqueeze(1), (1, 1, 1, 1), mode="replicate")
        q = torch.nn.functional.conv2d(q, k).squeeze(1)
        q = torch.floor(q + 0.5) + torch.randint(-2, 3, q.shape, generator=gen, device=q.device)
        dist.append(q.clamp(0, 255).to(torch.uint8).contiguous())
    return [p.contiguous() for p in ref], dist


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


# ----------------------------------------------------------------------------- CPU arm
def cpu_arm(sample_clip, yuv_main, yuv_ref, workers):
    """One bounded CPU sample of the same workload (all metrics at full resolution, I=1, plus
    PSNR/SSIM).  Returns (seconds, analysed frames, kind, engine description)."""
    from oracle import ref_port as RP
    engine = "cv2" if RP.cv2 is not None else "oracle"
    h, w = sample_clip.shape[1:3]
    t0 = time.perf_counter()
    RP.average_scene_complexity(sample_clip, w, h, frame_interval=1, workers=workers if engine == "cv2" else 1,
                                engine=engine)
    RP.psnr_ssim_frames(yuv_main, yuv_ref)
    dt = time.perf_counter() - t0
    desc = ("oracle/ref_port.py engine=cv2 (reference call structure: pool per metric, pickled frames, cv2 %s)"
            % RP.cv2.__version__) if engine == "cv2" else "oracle/ref_port.py engine=oracle (C + NumPy restatement, 1 thread)"
    return dt, len(sample_clip) - 1, "port", desc


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    from oracle import c_oracle
    c_oracle.build()
    import rtvqa_b200
    cores = os.cpu_count() or 1
    n = max(9, min(args.ref_frames, args.frames))
    clip = make_clip_host(n, args.height, args.width, seed=0)
    rng = np.random.default_rng(10_001)
    ry = np.empty((n, args.height, args.width), np.uint8)
    ru = np.empty((n, args.height // 2, args.width // 2), np.uint8)
    rv, dy, du, dv = np.empty_like(ru), np.empty_like(ry), np.empty_like(ru), np.empty_like(ru)
    for i in range(n):
        (a, b, c), (d, e, g) = rtvqa_b200.synth.synth_yuv_pair(clip[i], rng)
        ry[i], ru[i], rv[i], dy[i], du[i], dv[i] = a, b, c, d, e, g
    times = []
    for it in range(args.warmup + args.steps):
        dt, analysed, kind, desc = cpu_arm(clip, (dy, du, dv), (ry, ru, rv), cores)
        if it >= args.warmup:
            times.append(dt)
    per = float(np.mean(times))
    value = analysed / per
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/f32", "data": "synthetic",
        "config": {"workload": f"{args.width}x{args.height} synthetic clip, every frame, full-resolution, all 7 complexity "
                               f"metrics + PSNR/SSIM; CPU step = bounded sample of {n} frames ({analysed} analysed)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": f"{n} frames/step; {desc}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)
    return 0


# ----------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import rtvqa_b200
    from rtvqa_b200 import _native as N
    from rtvqa_b200 import complexity_metrics as cm
    from rtvqa_b200 import sharding as SH
    from rtvqa_b200 import video_processing as vp

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries the single JSON line: NCCL's own log (its "NCCL version ..." banner is printed at
        # every level from VERSION up, INFO traces when the operator asks for them) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    ctx = N.get_context(local)
    ctx.use_torch_stream()
    H, W, F = args.height, args.width, args.frames
    alpha = 0.8
    # default None = the reference's hard-wired 64x64 ORB (the headline workload); --orb-size WxH times the
    # full ORB pipeline of the orb_size knob instead and says so in `config`
    orb_size = tuple(int(v) for v in args.orb_size.lower().split("x")) if args.orb_size else None

    # ---- workload (outside every timed region) ------------------------------------------
    clip_host = torch.from_numpy(make_clip_host(F, H, W, seed=rank)).pin_memory()
    clip_dev = clip_host.to(dev, non_blocking=True)
    ref_dev, dist_dev = make_yuv_pairs_device(clip_dev, seed=rank)
    ref_host = [p.cpu().pin_memory() for p in ref_dev]
    dist_host = [p.cpu().pin_memory() for p in dist_dev]
    ts = rtvqa_b200.synth.synth_timestamps(F * world, 30.0)
    halo_dev = torch.empty((H, W, 3), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    k_total = F * world                                # sampled frames of the virtual stream (I = 1)
    a0 = rank * F

    def exchange_halo():
        """Previous rank's last frame -> this rank's halo (NCCL P2P over NVLink)."""
        if world == 1:
            return None
        ops = []
        if rank + 1 < world:
            ops.append(dist.P2POp(dist.isend, clip_dev[F - 1], rank + 1))
        if rank > 0:
            ops.append(dist.P2POp(dist.irecv, halo_dev, rank - 1))
        for w_ in dist.batch_isend_irecv(ops):
            w_.wait()
        return halo_dev if rank > 0 else None

    def step_device():
        halo = exchange_halo()
        rows = ctx.complexity_frames(clip_dev, W, H, N.M_ALL, halo=halo, orb_size=orb_size)
        fr = ctx.psnr_ssim(dist_dev, ref_dev)
        partials = SH.local_partials(rows, a0, k_total, alpha, ctx.ewm_partial)
        ints = np.array([int(rows["edge_count"].sum()), int(rows["orb_count"].sum()), len(rows)], dtype=np.int64)
        partials, ints = SH.reduce_partials(partials, ints)
        fps = ctx.framerate_series(ts)
        res = SH.finalize(partials, k_total, ctx.ewm_partial(fps, 0, len(fps), alpha))
        return res, fr, rows

    clip_np = clip_host.numpy()
    dist_np = [p.numpy() for p in dist_host]
    ref_np = [p.numpy() for p in ref_host]

    def step_e2e():
        """Public API with HOST buffers: the H2D copies and the D2H of the rows are inside.
        video_processing.analyze_frames = both halves with one interleaved upload schedule."""
        rows, fr = vp.analyze_frames(clip_np, W, H, dist_np, ref_np, local, orb_size=orb_size)
        vals = [cm._smoothed_mean(rows[name][SH.FIRST[name]:], alpha) for name in SH.SERIES]
        return vals, fr

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.kernel_launches()
        t0 = time.perf_counter()
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if world > 1:
            dist.barrier()
        ms = max(e0.elapsed_time(e1), 0.0)
        # the ABI calls end with a D2H + stream sync, so device time == wall time to within the
        # launch overhead; take the larger and the max over ranks
        sec = max(ms / 1e3, wall)
        t = torch.tensor([sec], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), ctx.kernel_launches() - l0, out

    sampler = ClockSampler(local)
    sampler.start()
    sec_dev, launches, out = timed(step_device, args.steps, args.warmup)
    clocks = sampler.summary()
    sec_e2e, _, out_e2e = timed(step_e2e, max(1, min(args.steps, 5)), 2)
    e2e_steps = max(1, min(args.steps, 5))

    analysed = F - 1 if world == 1 else F               # pairs per rank: rank 0 has no halo
    total_analysed = (F - 1) + (world - 1) * F
    value = total_analysed * args.steps / sec_dev
    e2e_value = (F - 1) * world * e2e_steps / sec_e2e

    # ---- roofline leg: per-kernel CUDA-event timing of one more pass (not part of `value`) ----
    ctx.kernel_profile(True)
    ctx.complexity_frames(clip_dev, W, H, N.M_ALL, orb_size=orb_size)
    ctx.psnr_ssim(dist_dev, ref_dev)
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
    top = max(rep.items(), key=lambda kv: kv[1]["ms"])
    tname, tv = top
    achieved = tv["bytes"] / (tv["ms"] * 1e-3) / 1e9 if tv["ms"] > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": tname, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": None, "peak_source": peak_src,
                "launches_per_step": tv["launches"], "avg_launch_ms": tv["ms"] / max(tv["launches"], 1),
                "share_of_step": tv["ms"] / tot_ms,
                "algorithmic_bytes_per_launch": tv["bytes"] / max(tv["launches"], 1),
                "kernels": {k: {"ms": round(v["ms"], 4), "launches": v["launches"],
                                "GBps": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else None,
                                "TFLOPs": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2) if v["flops"] else None}
                            for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"])},
                "end_to_end_input_GBps": value * 3 * H * W * 2 / 1e9 / world}
    traffic_file = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as f:
                t = json.load(f).get(tname)
            if t and t.get("ratio"):
                # DRAM bytes per launch = (dram bytes / algorithmic bytes of the ncu-captured level-0 launch)
                # x this run's algorithmic bytes per launch (launches differ in size across pyramid levels)
                roofline["traffic"] = t["ratio"] * roofline["algorithmic_bytes_per_launch"]
                roofline["traffic_source"] = {"report": t["report"], "captured_dram_bytes": t["dram_bytes"],
                                              "captured_algorithmic_bytes": t["algorithmic_bytes"], "ratio": t["ratio"]}
        except Exception:
            pass

    line = None
    if rank == 0:
        # ---- CPU baseline on the host cores, bounded sample (N = 1 only) ----
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import c_oracle
            c_oracle.build()
            n = max(9, min(args.ref_frames, F))
            cores = os.cpu_count() or 1
            dt, an, kind, desc = cpu_arm(clip_np[:n], [p[:n] for p in dist_np], [p[:n] for p in ref_np], cores)
            cpu = {"value": an / dt, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": f"first {n} frames of the bench clip ({an} analysed), {dt:.1f} s; {desc}"}
        res, fr, rows = out
        h2d = clip_np.nbytes + sum(p.nbytes for p in dist_np) + sum(p.nbytes for p in ref_np)
        d2h = rows.nbytes + fr.nbytes
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/f32 (integer pixel paths; fp32 flow/DCT, fp64 reductions)", "data": "synthetic",
            "config": {"workload": f"BASELINE.json configs[1]+[2]: {W}x{H} synthetic clip of {F} frames per GPU, every "
                                   f"frame (frame_interval 1), resize {W}x{H}, all 7 complexity metrics + PSNR/SSIM of {F} "
                                   "yuv420p pairs", "frames_per_gpu": F, "analysed_frames_per_step": total_analysed,
                       "l2": f"inputs ({(clip_np.nbytes + 2 * sum(p.nbytes for p in ref_np)) / 1e6:.0f} MB/step) larger than the 126 MB L2; no flush",
                       "parallelism": f"frame-range x{world}, one-frame halo over NCCL P2P, one all-reduce of partial sums"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "api": "video_processing.analyze_frames (vqa_analyze_clip) on pinned host arrays"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "result": {"scene_complexity": [float(v) for v in res], "psnr_avg_first": float(fr["psnr_avg"][0]),
                       "ssim_all_first": float(fr["ssim_all"][0])},
        }
        if orb_size:
            line["config"]["orb_size"] = "%dx%d (extension: full ORB pipeline instead of the reference's 64x64)" % orb_size
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
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
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--ref-frames", type=int, default=33, help="frames per CPU sample (bounded)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--orb-size", default="", help="WxH: full ORB pipeline on gray(resize(frame, WxH)) (default: reference 64x64)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
