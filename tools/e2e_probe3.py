import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rtvqa_b200
from rtvqa_b200 import _native as N, complexity_metrics as cm, video_processing as vp
import bench
F, H, W = 300, 1080, 1920
clip_host = torch.from_numpy(bench.make_clip_host(F, H, W, 0)).pin_memory()
clip_dev = clip_host.cuda()
ref_dev, dist_dev = bench.make_yuv_pairs_device(clip_dev, 0)
ref_np = [p.cpu().pin_memory().numpy() for p in ref_dev]; dist_np = [p.cpu().pin_memory().numpy() for p in dist_dev]
clip_np = clip_host.numpy()
ctx = N.get_context(0)
def T(fn, n=4):
    fn(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3
print("side", os.environ.get("VQA_SIDE_STREAM", "1"))
print("  complexity device ms", T(lambda: ctx.complexity_frames(clip_dev, W, H)))
print("  complexity host   ms", T(lambda: cm._clip_metrics(clip_np, W, H)))
print("  analyze_frames    ms", T(lambda: vp.analyze_frames(clip_np, W, H, dist_np, ref_np, 0)))
print("  complexity device ms", T(lambda: ctx.complexity_frames(clip_dev, W, H)))
