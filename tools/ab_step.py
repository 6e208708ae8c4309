"""A/B of run-time knobs on one resident clip: median device time of vqa_complexity_frames per variant.
usage: python tools/ab_step.py [frames] -- variants are (name, {env}) pairs read per call by the library."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rtvqa_b200
from rtvqa_b200 import _native as N

F = int(sys.argv[1]) if len(sys.argv) > 1 else 96
ctx = N.Context(0)
clip = torch.from_numpy(rtvqa_b200.synth.synth_clip(F, 1080, 1920, seed=0)).cuda()
KNOBS = ("VQA_MS_H", "VQA_CHUNK", "VQA_MAT_T4", "VQA_MAT_T4U")
SETS = {
    "knobs": [("default", {}), ("ms_h=64", {"VQA_MS_H": "64"}), ("ms_h=144", {"VQA_MS_H": "144"}),
              ("ms_h=192", {"VQA_MS_H": "192"}), ("ms_h=270", {"VQA_MS_H": "270"}), ("ms_h=540", {"VQA_MS_H": "540"}),
              ("chunk=24", {"VQA_CHUNK": "24"}), ("chunk=32", {"VQA_CHUNK": "32"}), ("chunk=96", {"VQA_CHUNK": "96"}),
              ("default again", {})],
    "mat_t4u": [("t4u=0", {"VQA_MAT_T4U": "0"}), ("t4u=1", {"VQA_MAT_T4U": "1"}), ("t4u=0 again", {"VQA_MAT_T4U": "0"}),
                ("t4u=1 again", {"VQA_MAT_T4U": "1"}), ("all old", {"VQA_MAT_T4U": "0", "VQA_MAT_T4": "0"})],
    "one": [("default", {}), ("default again", {})],
    "mat_t4": [("mat_t4=0", {"VQA_MAT_T4": "0"}), ("mat_t4=1", {"VQA_MAT_T4": "1"}), ("mat_t4=0 again", {"VQA_MAT_T4": "0"}),
               ("mat_t4=1 again", {"VQA_MAT_T4": "1"})],
}
VARIANTS = SETS[sys.argv[2] if len(sys.argv) > 2 else "knobs"]
base = None
for name, env in VARIANTS:
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update(env)
    ref = ctx.complexity_frames(clip, 1920, 1080)           # warm-up (allocations for this variant)
    ts = []
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rows = ctx.complexity_frames(clip, 1920, 1080)
        ts.append(time.perf_counter() - t0)
    if base is None:
        base = rows
    same = all(np.array_equal(rows[f], base[f], equal_nan=True) for f in rows.dtype.names)
    ctx.kernel_profile(True)
    ctx.complexity_frames(clip, 1920, 1080)
    rep = ctx.kernel_report()
    ctx.kernel_profile(False)
    blur = sum(v["ms"] for k, v in rep.items() if "blur_solve" in k and "EPI" not in k)
    mat = {k: round(v["ms"], 3) for k, v in rep.items() if "matrices" in k or (len(VARIANTS) <= 2 and v["ms"] > 0.25)}
    tot = sum(v["ms"] for v in rep.values())
    print(f"{name:14s} median {np.median(ts) * 1e3:8.2f} ms  min {min(ts) * 1e3:8.2f}  ({(F - 1) / np.median(ts):7.1f} frames/s)  "
          f"blur {blur:6.2f} ms  kernel-sum {tot:6.2f} ms  rows identical to default: {same}  {mat}", flush=True)
