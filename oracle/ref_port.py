"""Whole-clip CPU port of the reference's hot path, built on the oracle restatements.

TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's CPU arm).

Two engines:
  * ``engine="oracle"`` -- every operator from oracle/np_oracle.py + oracle/c/vqa_oracle.c
    (no OpenCV needed); single process.  This is the parity checker.
  * ``engine="cv2"``    -- the same call structure as the reference
    (complexity_metrics.py:246-310: one process pool per metric, frames pickled to workers,
    one cv2 call per frame) with the operators delegated to OpenCV exactly where the
    reference delegates them.  Used only as the timed CPU arm of bench.py when ``cv2`` is
    importable on the host (it is the reference's real cost model: pool start-up + pickling
    + cv2).  It is a re-implementation, not the reference's file.
"""
from __future__ import annotations

import functools
import os
from concurrent.futures import ProcessPoolExecutor

import numpy as np

from . import c_oracle as CO
from . import np_oracle as NO

try:  # optional: only the timed CPU arm wants it
    import cv2  # type: ignore
except Exception:  # pragma: no cover
    cv2 = None


# ------------------------------------------------------------------ per-item operators (oracle)
def o_motion(pair):
    cur, prev = pair
    if cur is None or prev is None:
        return 0.0
    return CO.farneback_mean_mag(NO.bgr2gray(prev), NO.bgr2gray(cur))


def o_dct(frame, rw, rh):
    return NO.process_dct_frame(frame, rw, rh)


def o_hist(frame, rw, rh):
    return NO.process_histogram_frame(frame, rw, rh)


def o_color(frame, rw, rh):
    return NO.process_color_histogram_frame(frame, rw, rh)


def o_edge(frame, rw, rh):
    return CO.canny_count(NO.bgr2gray(NO.resize_linear_u8(frame, rw, rh)))


def o_orb(frame):
    return CO.orb_count_64(NO.bgr2gray(NO.resize_linear_u8(frame, 64, 64)))


def o_tdct(prev_gray, cur_gray, rw, rh):
    return NO.process_temporal_dct_frame(prev_gray, cur_gray, rw, rh)


# ------------------------------------------------------------------ per-item operators (cv2)
def c_motion(pair):
    cur, prev = pair
    if cur is None or prev is None:
        return 0.0
    a = cv2.cvtColor(prev, cv2.COLOR_BGR2GRAY)
    b = cv2.cvtColor(cur, cv2.COLOR_BGR2GRAY)
    fl = cv2.calcOpticalFlowFarneback(a, b, None, 0.5, 3, 15, 3, 5, 1.2, 0)
    return np.mean(cv2.cartToPolar(fl[..., 0], fl[..., 1])[0])


def c_dct(frame, rw, rh):
    x = np.float32(cv2.resize(cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY), (rw, rh)))
    return np.sum(cv2.dct(x) ** 2)


def c_hist(frame, rw, rh):
    g = cv2.cvtColor(cv2.resize(frame, (rw, rh)), cv2.COLOR_BGR2GRAY)
    h = cv2.calcHist([g], [0], None, [256], [0, 256])
    p = h / h.sum()
    p = p[p > 0]
    return -np.sum(p * np.log2(p))


def c_color(frame, rw, rh):
    r = cv2.resize(frame, (rw, rh))
    acc = 0.0
    for ch in range(3):
        h = cv2.calcHist([r], [ch], None, [256], [0, 256])
        s = h.sum()
        if s == 0:
            return float("nan")
        p = h / s
        acc = acc + np.sum(p * np.log2(p + 1e-8))
    return -acc


def c_edge(frame, rw, rh):
    g = cv2.cvtColor(cv2.resize(frame, (rw, rh)), cv2.COLOR_BGR2GRAY)
    return np.sum(cv2.Canny(g, 100, 200) > 0)


def c_orb(frame):
    g = cv2.cvtColor(cv2.resize(frame, (64, 64)), cv2.COLOR_BGR2GRAY)
    return len(cv2.ORB_create().detectAndCompute(g, None)[0])


def c_tdct(prev_gray, cur_gray, rw, rh):
    a = cv2.dct(np.float32(cv2.resize(prev_gray, (rw, rh))))
    b = cv2.dct(np.float32(cv2.resize(cur_gray, (rw, rh))))
    return np.sum(np.abs(a - b))


# ------------------------------------------------------------------ executor
def pooled_map(items, fn, workers, batch=100):
    """Order-preserving chunked map; a fresh pool per call, like process_in_batches
    (complexity_metrics.py:128-148).  workers<=1 runs inline."""
    items = list(items)
    if workers is None or workers <= 1:
        return [fn(x) for x in items]
    out = []
    with ProcessPoolExecutor(max_workers=workers) as ex:
        for i in range(0, len(items), batch):
            out.extend(ex.map(fn, items[i:i + batch]))
    return out


def sample_pairs(clip, interval):
    idx = NO.sampled_indices(len(clip), interval)
    return [(clip[idx[j]], clip[idx[j - 1]]) for j in range(1, len(idx))]


def average_scene_complexity(clip, rw, rh, frame_interval=10, alpha=0.8, fps=30.0,
                             workers=1, batch=100, engine="oracle"):
    """calculate_average_scene_complexity (complexity_metrics.py:246-310) on an in-memory
    clip.  Returns the reference's 8-tuple order: motion, dct, hist, edge, orb, colour,
    temporal-dct, framerate."""
    if engine == "cv2":
        if cv2 is None:
            raise RuntimeError("cv2 engine requested but OpenCV is not importable")
        f_motion, f_dct, f_hist, f_color, f_edge, f_orb, f_tdct = c_motion, c_dct, c_hist, c_color, c_edge, c_orb, c_tdct
        gray = lambda f: cv2.resize(cv2.cvtColor(f, cv2.COLOR_BGR2GRAY), (rw, rh))
    else:
        f_motion, f_dct, f_hist, f_color, f_edge, f_orb, f_tdct = o_motion, o_dct, o_hist, o_color, o_edge, o_orb, o_tdct
        gray = lambda f: NO.resize_linear_u8(NO.bgr2gray(f), rw, rh)
    P = functools.partial
    pairs = sample_pairs(clip, frame_interval)
    frames = [p[0] for p in pairs]
    m = lambda series: NO.smoothed_mean(series, alpha)
    motion = m(pooled_map(pairs, f_motion, workers, batch))
    dct = m(pooled_map(frames, P(f_dct, rw=rw, rh=rh), workers, batch))
    hist = m(pooled_map(frames, P(f_hist, rw=rw, rh=rh), workers, batch))
    edge = m(pooled_map(frames, P(f_edge, rw=rw, rh=rh), workers, batch))
    orb = m(pooled_map(frames, f_orb, workers, batch))
    color = m(pooled_map(frames, P(f_color, rw=rw, rh=rh), workers, batch))
    # temporal DCT: serial in the parent, consecutive sampled grays of pair[0] (:524-541)
    td, prev_g = [], None
    for cur, _prev in pairs:
        g = gray(cur)
        if prev_g is not None:
            td.append(f_tdct(prev_g, g, rw, rh))
        prev_g = g
    tdct = NO.smoothed_mean(td, alpha) if td else 0.0
    ts = [1000.0 * i / fps for i in NO.timestamp_indices(len(clip), frame_interval)]
    fr = m(pooled_map(list(zip(ts[:-1], ts[1:])), NO.process_frame_interval_for_parallel, workers, batch))
    return (motion, dct, hist, edge, orb, color, tdct, fr)


def psnr_ssim_frames(main, ref):
    """Per-frame FFmpeg psnr/ssim on planar yuv420p stacks ((Y,U,V) of [n,h,w] uint8).
    Returns dict of arrays: mse[n,3], psnr_avg[n], ssim[n,3], ssim_all[n]."""
    n = main[0].shape[0]
    areas = np.array([p.shape[1] * p.shape[2] for p in main], dtype=np.float64)
    wgt = areas / areas.sum()
    mse = np.zeros((n, 3))
    ssim = np.zeros((n, 3))
    for i in range(n):
        for c in range(3):
            mse[i, c] = CO.plane_sse(main[c][i], ref[c][i]) / areas[c]
            ssim[i, c] = CO.ssim_plane(main[c][i], ref[c][i])
    mse_avg = mse @ wgt
    with np.errstate(divide="ignore"):
        psnr_avg = 10.0 * np.log10(255.0 * 255.0 / mse_avg)
    return dict(mse=mse, mse_avg=mse_avg, psnr_avg=psnr_avg, ssim=ssim, ssim_all=ssim @ wgt)
