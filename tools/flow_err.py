"""Mean-magnitude / per-pixel error of the CUDA Farneback flow against the C oracle.
usage: python tools/flow_err.py [h w]..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rtvqa_b200
from rtvqa_b200 import _native as N, synth
from oracle import c_oracle as CO, np_oracle as NO

ctx = N.get_context(0)
sizes = [(270, 480, 3), (540, 960, 5), (1080, 1920, 0)]
for h, w, seed in sizes:
    clip = synth.synth_clip(2, h, w, seed=seed)
    a, b = NO.bgr2gray(clip[0]), NO.bgr2gray(clip[1])
    want_mean, want_flow = CO.farneback_mean_mag(a, b, want_flow=True)
    flow = ctx.debug_flow(a, b)
    mag = np.sqrt(flow[..., 0].astype(np.float64) ** 2 + flow[..., 1] ** 2).mean()
    d = np.abs(flow - want_flow)
    print(f"{h}x{w}: mean |flow| {mag:.9f} oracle {float(want_mean):.9f} rel {abs(mag - want_mean) / want_mean:.2e}  "
          f"|d| median {np.median(d):.2e} p99 {np.percentile(d, 99):.2e} max {d.max():.2e}")
