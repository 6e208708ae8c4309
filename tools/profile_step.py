"""One hot-path step bracketed by cudaProfilerStart/Stop for `ncu --profile-from-start off`.
usage: python tools/profile_step.py [frames] [height] [width]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rtvqa_b200
from rtvqa_b200 import _native as N

F = int(sys.argv[1]) if len(sys.argv) > 1 else 24
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1920
ORB = tuple(int(v) for v in os.environ["VQA_PROF_ORB"].lower().split("x")) if os.environ.get("VQA_PROF_ORB") else None
ctx = N.Context(0)
if os.environ.get("VQA_PROF_BGR"):                         # round-1 entry: BGR frames + a separate PSNR/SSIM call
    clip = torch.from_numpy(rtvqa_b200.synth.synth_clip(F, H, W, seed=0)).cuda()
    (ry, ru, rv), (dy, du, dv) = rtvqa_b200.synth.synth_yuv_pairs(4, H, W, seed=1)
    step = lambda: (ctx.complexity_frames(clip, W, H, orb_size=ORB), ctx.psnr_ssim((dy, du, dv), (ry, ru, rv)))
else:                                                      # the bench's step: source + encode as yuv420p planes in HBM
    from rtvqa_b200.synth_device import DeviceClipSynth
    ref, enc = DeviceClipSynth(H, W, 0, torch.device("cuda", 0)).pairs(0, F)
    step = lambda: ctx.analyze_clip_yuv420(enc, ref, W, H, orb_size=ORB)
step()                                                     # warm-up (allocations, tensor maps, basis)
torch.cuda.synchronize()
torch.cuda.profiler.start()
rows, fr = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", rows["motion"][1:4], fr["psnr_avg"][:2])
if os.environ.get("VQA_PROF_REPORT"):                      # the library's own per-kernel table of the SAME step (declared bytes)
    import json
    ctx.kernel_profile(True)
    step()
    rep = ctx.kernel_report()
    ctx.kernel_profile(False)
    json.dump({"frames": F, "height": H, "width": W, "kernels": rep}, open(os.environ["VQA_PROF_REPORT"], "w"), indent=1)
