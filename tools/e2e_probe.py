import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rtvqa_b200
from rtvqa_b200 import _native as N, complexity_metrics as cm, video_processing as vp
from concurrent.futures import ThreadPoolExecutor
import bench
F, H, W = 300, 1080, 1920
clip_host = torch.from_numpy(bench.make_clip_host(F, H, W, 0)).pin_memory()
clip_dev = clip_host.cuda()
ref_dev, dist_dev = bench.make_yuv_pairs_device(clip_dev, 0)
ref_np = [p.cpu().pin_memory().numpy() for p in ref_dev]; dist_np = [p.cpu().pin_memory().numpy() for p in dist_dev]
clip_np = clip_host.numpy()
ctx = N.get_context(0)
def T(fn, n=3):
    fn(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3
print("complexity device ms", T(lambda: ctx.complexity_frames(clip_dev, W, H)))
print("complexity host   ms", T(lambda: cm._clip_metrics(clip_np, W, H)))
print("psnr device ms", T(lambda: ctx.psnr_ssim(dist_dev, ref_dev)))
print("psnr host   ms", T(lambda: vp.psnr_ssim_frames(dist_np, ref_np, 0)))
pool = ThreadPoolExecutor(1)
def both():
    f = pool.submit(vp.psnr_ssim_frames, dist_np, ref_np, 0); cm._clip_metrics(clip_np, W, H); f.result()
print("both threaded ms", T(both))
for ch in (8, 16, 24):
    os.environ["VQA_CHUNK"] = str(ch)
    print("chunk", ch, "complexity host ms", T(lambda: cm._clip_metrics(clip_np, W, H)), "device ms", T(lambda: ctx.complexity_frames(clip_dev, W, H)))
import threading
os.environ.pop("VQA_CHUNK", None)
for frc in (64, 8):
    os.environ["VQA_FR_CHUNK"] = str(frc)
    print("FR chunk", frc, "psnr host ms", T(lambda: vp.psnr_ssim_frames(dist_np, ref_np, 0)), "both threaded ms", T(both))
def both_delayed(delay):
    def run():
        time.sleep(delay); return vp.psnr_ssim_frames(dist_np, ref_np, 0)
    f = pool.submit(run); cm._clip_metrics(clip_np, W, H); f.result()
for d in (0.03, 0.06, 0.09):
    print("delayed", d, T(lambda: both_delayed(d)))
def seq():
    cm._clip_metrics(clip_np, W, H); vp.psnr_ssim_frames(dist_np, ref_np, 0)
print("sequential ms", T(seq))
