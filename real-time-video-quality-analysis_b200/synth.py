"""Deterministic synthetic clips (NumPy only) shared by tests, the oracle and bench.py.

Recipe = SURVEY.md §8(d): a blurred-noise background panned 3 px/frame in x and 2 px/frame
in y, a moving filled white circle, a moving filled red rectangle and 0..7 of additive
per-pixel noise.  Frames are uint8 HWC in B,G,R order, exactly what
``cv2.VideoCapture.read`` hands the reference (complexity_metrics.py:99-106).

The generator is a *workload* definition, not part of the measured path.
"""
from __future__ import annotations

import numpy as np

__all__ = ["synth_clip", "synth_frame_iter", "synth_timestamps", "bgr_to_yuv420",
           "synth_yuv_pair", "synth_yuv_pairs"]


def _gauss_kernel(sigma: float) -> np.ndarray:
    r = int(np.ceil(3.0 * sigma))
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-(x * x) / (2.0 * sigma * sigma))
    return (k / k.sum()).astype(np.float32)


def _blur_axis(a: np.ndarray, k: np.ndarray, axis: int) -> np.ndarray:
    """Reflect-101 separable convolution along ``axis`` (float32)."""
    r = len(k) // 2
    pad = [(0, 0)] * a.ndim
    pad[axis] = (r, r)
    p = np.pad(a, pad, mode="reflect")
    out = np.zeros_like(a, dtype=np.float32)
    n = a.shape[axis]
    for i, kv in enumerate(k):
        sl = [slice(None)] * a.ndim
        sl[axis] = slice(i, i + n)
        out += kv * p[tuple(sl)]
    return out


def _background(n: int, h: int, w: int, rng: np.random.Generator) -> np.ndarray:
    raw = rng.integers(0, 256, (h + 2 * n, w + 3 * n, 3), dtype=np.uint8).astype(np.float32)
    k = _gauss_kernel(2.0)
    b = _blur_axis(_blur_axis(raw, k, 0), k, 1)
    lo, hi = float(b.min()), float(b.max())
    b = (b - lo) * (255.0 / max(hi - lo, 1e-6))
    return np.clip(np.rint(b), 0, 255).astype(np.uint8)


def synth_frame_iter(n: int, h: int, w: int, seed: int = 0):
    """Yield ``n`` BGR uint8 frames of ``h x w`` one at a time (bounded host memory)."""
    rng = np.random.default_rng(seed)
    bg = _background(n, h, w, rng)
    yy, xx = np.mgrid[0:h, 0:w]
    rad2 = (h / 12.0) ** 2
    for i in range(n):
        f = bg[2 * i:2 * i + h, 3 * i:3 * i + w].copy()
        cx, cy = w / 4.0 + 5 * i, h / 3.0 + 2 * i
        f[(xx - cx) ** 2 + (yy - cy) ** 2 <= rad2] = 255
        x0, y0 = int(w / 2), int(h / 5 + 3 * i)
        x1, y1 = int(w / 2 + w / 6), int(h / 5 + h / 4 + 3 * i)
        f[max(y0, 0):max(min(y1, h), 0), x0:min(x1, w)] = (0, 0, 255)
        noise = rng.integers(0, 8, (h, w, 3), dtype=np.uint8)
        f = np.minimum(f.astype(np.uint16) + noise, 255).astype(np.uint8)
        yield np.ascontiguousarray(f)


def synth_clip(n: int, h: int, w: int, seed: int = 0) -> np.ndarray:
    """``(n, h, w, 3)`` uint8 BGR clip."""
    out = np.empty((n, h, w, 3), dtype=np.uint8)
    for i, f in enumerate(synth_frame_iter(n, h, w, seed)):
        out[i] = f
    return out


def synth_timestamps(n: int, fps: float = 30.0) -> np.ndarray:
    """CAP_PROP_POS_MSEC of a CFR clip: ``1000*i/fps``."""
    return 1000.0 * np.arange(n, dtype=np.float64) / float(fps)


def bgr_to_yuv420(frame: np.ndarray):
    """BT.601 limited-range BGR -> planar 4:2:0 (Y HxW, U/V H/2 x W/2), uint8.

    Only a workload generator for the PSNR/SSIM half (stands in for the decoded yuv420p
    planes FFmpeg's filters see); not required to match any library bit for bit.
    """
    b = frame[..., 0].astype(np.int32)
    g = frame[..., 1].astype(np.int32)
    r = frame[..., 2].astype(np.int32)
    y = (66 * r + 129 * g + 25 * b + 128 >> 8) + 16
    u = (-38 * r - 74 * g + 112 * b + 128 >> 8) + 128
    v = (112 * r - 94 * g - 18 * b + 128 >> 8) + 128
    h2, w2 = frame.shape[0] // 2 * 2, frame.shape[1] // 2 * 2

    def sub(p):
        p = p[:h2, :w2]
        return (p[0::2, 0::2] + p[0::2, 1::2] + p[1::2, 0::2] + p[1::2, 1::2] + 2) >> 2

    return (np.clip(y, 0, 255).astype(np.uint8),
            np.clip(sub(u), 0, 255).astype(np.uint8),
            np.clip(sub(v), 0, 255).astype(np.uint8))


def _blur3(p: np.ndarray) -> np.ndarray:
    q = np.pad(p.astype(np.int32), 1, mode="edge")
    hsum = q[:, :-2] + 2 * q[:, 1:-1] + q[:, 2:]
    return (hsum[:-2] + 2 * hsum[1:-1] + hsum[2:] + 8) >> 4


def synth_yuv_pair(frame: np.ndarray, rng: np.random.Generator):
    """(reference planes, distorted planes): distortion = 3x3 blur + integers(-2,3) noise,
    a stand-in for the CRF-23 re-encode of BASELINE.json config 3 (no x264 in the image)."""
    ref = bgr_to_yuv420(frame)
    dist = []
    for p in ref:
        d = _blur3(p) + rng.integers(-2, 3, p.shape)
        dist.append(np.clip(d, 0, 255).astype(np.uint8))
    return ref, tuple(dist)


def synth_yuv_pairs(n: int, h: int, w: int, seed: int = 1):
    """Planar stacks: ref (Y[n,h,w], U[n,h/2,w/2], V) and dist, uint8."""
    rng = np.random.default_rng(10_000 + seed)
    ry = np.empty((n, h, w), np.uint8)
    ru = np.empty((n, h // 2, w // 2), np.uint8)
    rv = np.empty_like(ru)
    dy, du, dv = np.empty_like(ry), np.empty_like(ru), np.empty_like(ru)
    for i, f in enumerate(synth_frame_iter(n, h, w, seed)):
        (a, b, c), (d, e, g) = synth_yuv_pair(f, rng)
        ry[i], ru[i], rv[i], dy[i], du[i], dv[i] = a, b, c, d, e, g
    return (ry, ru, rv), (dy, du, dv)
