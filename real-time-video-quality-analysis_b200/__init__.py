"""B200-native per-frame scene-complexity + PSNR/SSIM path (drop-in for the hot path of
zaki699/Real-Time-Video-Quality-Analysis: complexity_metrics.py + run_ffmpeg_metrics).

Layout:
  csrc/                 hand-written sm_100a CUDA kernels + the C ABI (include/vqa_b200.h)
  _native.py            ctypes binding of libvqa_b200.so (fails loudly if it is missing)
  complexity_metrics.py host mirror of the reference module (same names / signatures)
  video_processing.py   host mirror of the reference CLI module (PSNR/SSIM on the GPU)
  sharding.py           frame-range sharding + NCCL reduce (one process per GPU)
  synth.py              deterministic synthetic clips (workload generator)
"""
from . import synth  # noqa: F401

__all__ = ["synth"]
__version__ = "0.1.0"
