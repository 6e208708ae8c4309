"""ctypes loader for the plain-C oracle (oracle/c/vqa_oracle.c -> oracle/_ref/libvqa_oracle.so).

TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py CPU arm).  Build with ``make -C oracle``
(``__graft_entry__.build()`` does it)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libvqa_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "c", "vqa_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        u8p, f32p = C.POINTER(C.c_uint8), C.POINTER(C.c_float)
        L.vqo_bgr2gray.argtypes = [u8p, C.c_int, C.c_int, u8p]
        L.vqo_resize_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int]
        L.vqo_canny_count.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
        L.vqo_canny_count.restype = C.c_long
        L.vqo_fast_count.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
        L.vqo_fast_count.restype = C.c_int
        L.vqo_orb_count_64.argtypes = [u8p]
        L.vqo_orb_count_64.restype = C.c_int
        L.vqo_farneback_mean_mag.argtypes = [u8p, u8p, C.c_int, C.c_int, f32p]
        L.vqo_farneback_mean_mag.restype = C.c_float
        L.vqo_plane_sse.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.vqo_plane_sse.restype = C.c_uint64
        L.vqo_ssim_plane.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.vqo_ssim_plane.restype = C.c_double
        L.vqo_polyexp_setup.argtypes = [C.c_int, C.c_double, f32p, f32p, f32p, C.POINTER(C.c_double)]
        _lib = L
    return _lib


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint8))


def bgr2gray(frame):
    f, fp = _u8(frame)
    out = np.empty(f.shape[:2], np.uint8)
    lib().vqo_bgr2gray(fp, f.shape[0], f.shape[1], out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out


def resize_u8(img, dw, dh):
    s, sp = _u8(img)
    cn = 1 if s.ndim == 2 else s.shape[2]
    out = np.empty((dh, dw) if s.ndim == 2 else (dh, dw, cn), np.uint8)
    lib().vqo_resize_u8(sp, s.shape[0], s.shape[1], cn, out.ctypes.data_as(C.POINTER(C.c_uint8)), dh, dw)
    return out


def canny_count(gray, low=100, high=200, want_map=False):
    """cv2.Canny(gray, low, high) edge-pixel count (complexity_metrics.py:503-504)."""
    g, gp = _u8(gray)
    m = np.empty_like(g) if want_map else None
    n = lib().vqo_canny_count(gp, g.shape[0], g.shape[1], low, high,
                              m.ctypes.data_as(C.POINTER(C.c_uint8)) if want_map else None)
    return (np.int64(n), m) if want_map else np.int64(n)


def fast_count(gray, thr=20, border=31, want_map=False):
    g, gp = _u8(gray)
    m = np.empty_like(g) if want_map else None
    n = lib().vqo_fast_count(gp, g.shape[0], g.shape[1], thr, border,
                             m.ctypes.data_as(C.POINTER(C.c_uint8)) if want_map else None)
    return (int(n), m) if want_map else int(n)


def orb_count_64(gray64):
    """len(cv2.ORB_create().detectAndCompute(gray64)) for a 64x64 image (complexity_metrics.py:385-387)."""
    g, gp = _u8(gray64)
    assert g.shape == (64, 64)
    return int(lib().vqo_orb_count_64(gp))


def farneback_mean_mag(prev_gray, next_gray, want_flow=False):
    """np.mean(|calcOpticalFlowFarneback(prev, next, None, .5, 3, 15, 3, 5, 1.2, 0)|)
    (complexity_metrics.py:340-343)."""
    p, pp = _u8(prev_gray)
    n, nq = _u8(next_gray)
    assert p.shape == n.shape
    flow = np.empty(p.shape + (2,), np.float32) if want_flow else None
    v = lib().vqo_farneback_mean_mag(pp, nq, p.shape[0], p.shape[1],
                                     flow.ctypes.data_as(C.POINTER(C.c_float)) if want_flow else None)
    return (np.float32(v), flow) if want_flow else np.float32(v)


def plane_sse(a, b):
    a, ap = _u8(a)
    b, bp = _u8(b)
    return int(lib().vqo_plane_sse(ap, bp, a.shape[0], a.shape[1], a.shape[1], b.shape[1]))


def ssim_plane(a, b):
    a, ap = _u8(a)
    b, bp = _u8(b)
    return float(lib().vqo_ssim_plane(ap, bp, a.shape[0], a.shape[1], a.shape[1], b.shape[1]))


def polyexp_setup(n=5, sigma=1.2):
    g = np.empty(2 * n + 1, np.float32)
    xg, xxg = np.empty_like(g), np.empty_like(g)
    ig = np.empty(4, np.float64)
    f32p = C.POINTER(C.c_float)
    lib().vqo_polyexp_setup(n, sigma, g.ctypes.data_as(f32p), xg.ctypes.data_as(f32p),
                            xxg.ctypes.data_as(f32p), ig.ctypes.data_as(C.POINTER(C.c_double)))
    return g, xg, xxg, ig
