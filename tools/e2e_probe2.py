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
fr = N.get_context(0, role="fr")
pool = ThreadPoolExecutor(1)
def timed_fr(host):
    t = time.perf_counter()
    (fr.psnr_ssim(dist_np, ref_np) if host else fr.psnr_ssim(dist_dev, ref_dev))
    return (time.perf_counter() - t) * 1e3
for host in (True, False):
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        f = pool.submit(timed_fr, host)
        t1 = time.perf_counter(); cm._clip_metrics(clip_np, W, H); tc = (time.perf_counter() - t1) * 1e3
        tf = f.result(); tot = (time.perf_counter() - t0) * 1e3
        print(f"FR host={host}: complexity {tc:.1f} ms, FR {tf:.1f} ms, total {tot:.1f} ms")
