"""Per-step device time of the bench workload over many consecutive steps (diagnostic: warm-up /
box variance).  usage: python tools/step_series.py [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, rtvqa_b200
from rtvqa_b200 import _native as N

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dev = torch.device("cuda", 0)
ctx = N.get_context(0)
ctx.use_torch_stream()
F, H, W = 300, 1080, 1920
clip = torch.from_numpy(bench.make_clip_host(F, H, W, seed=0)).to(dev)
ref_dev, dist_dev = bench.make_yuv_pairs_device(clip, seed=0)
torch.cuda.synchronize()
import pynvml
pynvml.nvmlInit()
hd = pynvml.nvmlDeviceGetHandleByIndex(0)
out = []
for i in range(steps):
    t0 = time.perf_counter()
    ctx.complexity_frames(clip, W, H, N.M_ALL)
    t1 = time.perf_counter()
    ctx.psnr_ssim(dist_dev, ref_dev)
    t2 = time.perf_counter()
    out.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, pynvml.nvmlDeviceGetClockInfo(hd, pynvml.NVML_CLOCK_SM),
                pynvml.nvmlDeviceGetClockInfo(hd, pynvml.NVML_CLOCK_MEM), pynvml.nvmlDeviceGetPowerUsage(hd) / 1000,
                pynvml.nvmlDeviceGetTemperature(hd, 0)))
for i, r in enumerate(out):
    print(f"step {i:3d}: complexity {r[0]:7.2f} ms  psnr/ssim {r[1]:6.2f} ms  sm {r[2]} MHz mem {r[3]} MHz  {r[4]:.0f} W  {r[5]} C")
