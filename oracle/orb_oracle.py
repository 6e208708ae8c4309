"""TEST INFRASTRUCTURE (CPU oracle) -- general-size ORB keypoint detection and scoring.

Restates what ``cv2.ORB_create().detectAndCompute(gray, None)`` does up to the keypoint list
(complexity_metrics.py:385-387 calls it on a 64x64 image; the ``orb_size`` knob of this repo lifts the
hard-wired size).  The arithmetic lives in opencv-python 4.10.0.84 (requirements.txt:2), un-vendored:
modules/features2d/src/orb.cpp (computeKeyPoints, HarrisResponses), fast.cpp / fast_score.cpp,
keypoint.cpp (KeyPointsFilter::retainBest, runByImageBorder) and imgproc/src/resize.cpp
(resize_bitExact, INTER_LINEAR_EXACT).  Pinned against cv2 4.13.0 of this image by
oracle/make_golden.py (tests/golden/orb_general.json) and tests/test_oracle_golden.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import numpy as np

from . import c_oracle

NFEATURES, SCALE_FACTOR, NLEVELS, EDGE_THRESHOLD, FAST_THRESHOLD = 500, 1.2, 8, 31, 20
HARRIS_BLOCK, HARRIS_K = 7, np.float32(0.04)


# ----------------------------------------------------------------------------- INTER_LINEAR_EXACT
def linear_exact_taps(sn: int, dn: int):
    """resize.cpp interpolationLinear<ufixedpoint16>::getCoeffs: per destination index the source
    offset and the two 8.8 fixed-point weights.  All arithmetic in IEEE double (softdouble), weights
    rounded half-to-even (cvRound).  Outside [minofst, maxofst) the edge pixel is replicated."""
    scale = 1.0 / (dn / sn)
    off = np.zeros(dn, np.int64)
    c0 = np.full(dn, 256, np.int64)
    c1 = np.zeros(dn, np.int64)
    for d in range(dn):
        f = scale * (d + 0.5) - 0.5
        i = int(np.floor(f))
        if i >= 0 and sn > 1:
            if i < sn - 1:
                off[d] = i
                c1[d] = int(np.rint((f - i) * 256.0))
                c0[d] = 256 - c1[d]
            else:
                off[d] = sn - 1
        # else: offset 0, weights (256, 0)
    return off, c0, c1


def resize_linear_exact_u8(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT) for one-channel uint8."""
    sh, sw = img.shape
    xo, xa, xb = linear_exact_taps(sw, dw)
    yo, ya, yb = linear_exact_taps(sh, dh)
    s = img.astype(np.int64)
    x1 = np.minimum(xo + 1, sw - 1)
    y1 = np.minimum(yo + 1, sh - 1)
    hl = s[:, xo] * xa[None, :] + s[:, x1] * xb[None, :]          # 8.8 fixed point, <= 255*256
    v = hl[yo, :] * ya[:, None] + hl[y1, :] * yb[:, None]         # 16.16
    return ((v + (1 << 15)) >> 16).astype(np.uint8)


# ----------------------------------------------------------------------------- pyramid geometry
def level_scales(nlevels: int = NLEVELS, scale_factor: float = SCALE_FACTOR):
    """orb.cpp getScale: (float)pow((double)(float)scaleFactor, level)."""
    sf = float(np.float32(scale_factor))
    return [np.float32(sf ** l) for l in range(nlevels)]


def level_sizes(h: int, w: int, nlevels: int = NLEVELS, scale_factor: float = SCALE_FACTOR):
    """cvRound(cols/scale), cvRound(rows/scale) in float32 (orb.cpp detectAndCompute)."""
    out = []
    for s in level_scales(nlevels, scale_factor):
        out.append((int(np.rint(np.float32(h) / s)), int(np.rint(np.float32(w) / s))))
    return out


def level_quotas(nfeatures: int = NFEATURES, nlevels: int = NLEVELS, scale_factor: float = SCALE_FACTOR):
    """orb.cpp computeKeyPoints nfeaturesPerLevel: float32 geometric series, cvRound per level, the
    remainder to the last level.  [109, 90, 75, 63, 52, 44, 36, 31] for the defaults."""
    f32 = np.float32
    factor = f32(1.0 / float(f32(scale_factor)))
    nd = f32(nfeatures) * (f32(1) - factor) / (f32(1) - f32(float(factor) ** nlevels))
    q, total = [], 0
    for _ in range(nlevels - 1):
        q.append(int(np.rint(nd)))
        total += q[-1]
        nd = f32(nd * factor)
    q.append(max(nfeatures - total, 0))
    return q


def pyramid(gray: np.ndarray, nlevels: int = NLEVELS, scale_factor: float = SCALE_FACTOR):
    levels = [np.ascontiguousarray(gray)]
    for (lh, lw) in level_sizes(gray.shape[0], gray.shape[1], nlevels, scale_factor)[1:]:
        levels.append(resize_linear_exact_u8(levels[-1], lw, lh))
    return levels


# ----------------------------------------------------------------------------- scoring and selection
def harris_response(img: np.ndarray, x: int, y: int) -> np.float32:
    """orb.cpp HarrisResponses: 7x7 block of Sobel-like integer gradients around (x, y); float32
    expression evaluated left to right without contraction."""
    r = HARRIS_BLOCK // 2
    p = img.astype(np.int64)

    def sh(dy, dx):
        return p[y - r + dy:y + r + 1 + dy, x - r + dx:x + r + 1 + dx]

    ix = (sh(0, 1) - sh(0, -1)) * 2 + (sh(-1, 1) - sh(-1, -1)) + (sh(1, 1) - sh(1, -1))
    iy = (sh(1, 0) - sh(-1, 0)) * 2 + (sh(1, -1) - sh(-1, -1)) + (sh(1, 1) - sh(-1, 1))
    a, b, c = np.float32(int((ix * ix).sum())), np.float32(int((iy * iy).sum())), np.float32(int((ix * iy).sum()))
    scale = np.float32(1.0) / (np.float32(4 * HARRIS_BLOCK) * np.float32(255.0))
    s4 = scale * scale * scale * scale
    ab = a + b
    return np.float32((a * b - c * c - HARRIS_K * ab * ab) * s4)


def retain_best(responses: np.ndarray, n: int) -> np.ndarray:
    """keypoint.cpp KeyPointsFilter::retainBest: boolean keep-mask; every item whose response is >= the
    n-th best is kept (ties at the boundary survive)."""
    m = len(responses)
    if n < 0 or m <= n:
        return np.ones(m, bool)
    if n == 0:
        return np.zeros(m, bool)
    nth = np.sort(responses)[::-1][n - 1]
    return responses >= nth


def orb_detect(gray: np.ndarray, nfeatures: int = NFEATURES, nlevels: int = NLEVELS,
               scale_factor: float = SCALE_FACTOR, edge_threshold: int = EDGE_THRESHOLD,
               fast_threshold: int = FAST_THRESHOLD):
    """Keypoints of cv2.ORB_create(nfeatures, scale_factor, nlevels, edge_threshold) as rows
    (level, x, y, harris_response, fast_score) in level coordinates, plus the per-level counts."""
    quotas = level_quotas(nfeatures, nlevels, scale_factor)
    rows, per_level = [], []
    for l, img in enumerate(pyramid(gray, nlevels, scale_factor)):
        lh, lw = img.shape
        kept = 0
        if lw > 2 * edge_threshold and lh > 2 * edge_threshold and min(lh, lw) >= 7:
            _, kmap = c_oracle.fast_count(img, fast_threshold, edge_threshold, want_map=True)
            ys, xs = np.nonzero(kmap)
            score = kmap[ys, xs].astype(np.float32)
            k1 = retain_best(score, 2 * quotas[l])
            ys, xs, score = ys[k1], xs[k1], score[k1]
            resp = np.array([harris_response(img, int(x), int(y)) for x, y in zip(xs, ys)], np.float32)
            k2 = retain_best(resp, quotas[l])
            for x, y, r, s in zip(xs[k2], ys[k2], resp[k2], score[k2]):
                rows.append((l, int(x), int(y), float(r), int(s)))
            kept = int(k2.sum())
        per_level.append(kept)
    return rows, per_level


# ----------------------------------------------------------------------------- orientation (ICAngles)
PATCH_SIZE = 31
HALF_PATCH = PATCH_SIZE // 2


def umax_table(half: int = HALF_PATCH):
    """orb.cpp computeKeyPoints: half-width of the circular patch per row offset v,
    [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3] for patchSize 31."""
    umax = [0] * (half + 2)
    vmax = int(np.floor(half * np.sqrt(2.0) / 2 + 1))
    vmin = int(np.ceil(half * np.sqrt(2.0) / 2))
    for v in range(vmax + 1):
        umax[v] = int(np.rint(np.sqrt(float(half * half - v * v))))
    v0 = 0
    for v in range(half, vmin - 1, -1):
        while umax[v0] == umax[v0 + 1]:
            v0 += 1
        umax[v] = v0
        v0 += 1
    return umax[:half + 1]


_F = np.float32
_RAD = _F(180.0 / np.pi)
ATAN_P1, ATAN_P3 = _F(0.9997878412794807) * _RAD, _F(-0.3258083974640975) * _RAD
ATAN_P5, ATAN_P7 = _F(0.1555786518463281) * _RAD, _F(-0.04432655554792128) * _RAD
ATAN_EPS = _F(2.220446049250313e-16)


def fast_atan2(y, x) -> np.float32:
    """core/mathfuncs_core fastAtan2 (degrees, odd degree-7 polynomial), float32, no contraction."""
    y, x = _F(y), _F(x)
    ax, ay = abs(x), abs(y)
    if ax >= ay:
        c = ay / (ax + ATAN_EPS)
        c2 = c * c
        a = (((ATAN_P7 * c2 + ATAN_P5) * c2 + ATAN_P3) * c2 + ATAN_P1) * c
    else:
        c = ax / (ay + ATAN_EPS)
        c2 = c * c
        a = _F(90.0) - (((ATAN_P7 * c2 + ATAN_P5) * c2 + ATAN_P3) * c2 + ATAN_P1) * c
    if x < 0:
        a = _F(180.0) - a
    if y < 0:
        a = _F(360.0) - a
    return _F(a)


def _reflect101(i: int, n: int) -> int:
    while i < 0 or i >= n:
        i = -i if i < 0 else 2 * n - 2 - i
    return i


_PATCH_CACHE = {}


def _patch_weights(umax):
    """(U * mask, V * mask) of the circular patch: mask[v, u] = |u| <= umax[|v|]."""
    key = tuple(umax)
    if key not in _PATCH_CACHE:
        half = len(umax) - 1
        v, u = np.mgrid[-half:half + 1, -half:half + 1]
        mask = np.abs(u) <= np.asarray(umax)[np.abs(v)]
        _PATCH_CACHE[key] = ((u * mask).astype(np.int64), (v * mask).astype(np.int64))
    return _PATCH_CACHE[key]


def ic_angle(img: np.ndarray, x: int, y: int, umax=None, padded=None) -> np.float32:
    """orb.cpp ICAngles: intensity-centroid orientation over the circular patch of radius 15
    (m10 = sum u*I, m01 = sum v*I, integers); pixels outside the level image come from its
    BORDER_REFLECT_101 extension (only reachable when edgeThreshold < 16).  ``padded`` = the level
    reflect-padded by the patch radius (pass it when calling for many keypoints)."""
    umax = umax or umax_table()
    half = len(umax) - 1
    if padded is None:
        padded = np.pad(img, half, mode="reflect")
    wu, wv = _patch_weights(umax)
    patch = padded[y:y + 2 * half + 1, x:x + 2 * half + 1].astype(np.int64)
    return fast_atan2(_F(int((patch * wv).sum())), _F(int((patch * wu).sum())))


def orb_keypoints(gray: np.ndarray, **kw):
    """cv2.KeyPoint fields of ORB's keypoints as rows (octave, pt.x, pt.y, size, angle, response): level
    coordinates times the float32 level scale, size = 31 * scale, ICAngles orientation."""
    nlevels, sf = kw.get("nlevels", NLEVELS), kw.get("scale_factor", SCALE_FACTOR)
    rows, _ = orb_detect(gray, **kw)
    pyr, scales, um = pyramid(gray, nlevels, sf), level_scales(nlevels, sf), umax_table()
    padded = {}
    out = []
    for l, x, y, r, _s in rows:
        sc = scales[l]
        if l not in padded:
            padded[l] = np.pad(pyr[l], HALF_PATCH, mode="reflect")
        out.append((l, float(_F(x) * sc), float(_F(y) * sc), float(_F(PATCH_SIZE) * sc),
                    float(ic_angle(pyr[l], x, y, um, padded[l])), r))
    return out


def orb_count(gray: np.ndarray, **kw) -> int:
    """len(cv2.ORB_create(**kw).detectAndCompute(gray, None)[0])"""
    return sum(orb_detect(gray, **kw)[1])
