"""CPU oracle (NumPy) for the per-frame complexity + PSNR/SSIM hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs use it, and only as the checker / the timed CPU arm.

The reference (zaki699/Real-Time-Video-Quality-Analysis) holds no arithmetic of its own for
this path: every operator is a call into un-vendored third parties --
opencv-python 4.10.0.84, pandas 2.2.3, numpy 2.1.2 (requirements.txt:1-3) and the FFmpeg CLI
filters ``psnr``/``ssim`` (video_processing.py:274-291).  This file restates the *published
algorithms* of those calls; each function cites the reference call site it stands in for.

Pinning status: the reference ships no tests or golden vectors, so the
restatement is pinned against the reference *itself*, imported from /root/reference in the
build container with cv2 4.13.0 (``oracle/make_golden.py`` -> ``tests/golden/*.npz``;
``tests/test_oracle_golden.py``).  PSNR/SSIM cannot be pinned that way (no ffmpeg binary in
the image): that half is "parity unpinned" beyond the known-answer cases of SURVEY.md A.9.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------- a1: ingest


def bgr2gray(frame: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(frame, COLOR_BGR2GRAY) for uint8 (complexity_metrics.py:327,358,405,493).

    15-bit fixed point: (3735 B + 19235 G + 9798 R + 2^14) >> 15   (SURVEY.md A.1)."""
    f = frame.astype(np.int32)
    return ((3735 * f[..., 0] + 19235 * f[..., 1] + 9798 * f[..., 2] + (1 << 14)) >> 15).astype(np.uint8)


def _linear_taps(sn: int, dn: int, vertical: bool = False):
    """Index/weight tables of cv2.resize INTER_LINEAR for one axis.

    Horizontal taps clamp the fraction to 0 at the borders; vertical taps keep the fraction
    and clip the two row indices instead (OpenCV resizeGeneric_) -- only distinguishable when
    upscaling."""
    d = np.arange(dn, dtype=np.float64)
    f = ((d + 0.5) * (float(sn) / float(dn)) - 0.5).astype(np.float32)
    i = np.floor(f).astype(np.int64)
    a = (f - i.astype(np.float32)).astype(np.float32)
    if not vertical:
        lo = i < 0
        i[lo] = 0
        a[lo] = 0.0
        hi = i >= sn - 1
        i[hi] = sn - 1
        a[hi] = 0.0
    i1 = np.clip(i + 1, 0, sn - 1)
    i = np.clip(i, 0, sn - 1)
    w1 = np.rint(a * np.float32(2048.0)).astype(np.int32)
    w0 = np.rint((np.float32(1.0) - a) * np.float32(2048.0)).astype(np.int32)
    return i, i1, w0, w1


def resize_linear_u8(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(img, (dw, dh)) default INTER_LINEAR on uint8, 1 or 3 channels
    (complexity_metrics.py:359,386,404,430,490,531).  No antialiasing; identity when the size
    is unchanged."""
    sh, sw = img.shape[:2]
    if (sw, sh) == (dw, dh):
        return img.copy()
    x0, x1, a0, a1 = _linear_taps(sw, dw)
    y0, y1, b0, b1 = _linear_taps(sh, dh, vertical=True)
    s = img.astype(np.int32)
    if s.ndim == 3:
        a0_, a1_ = a0[None, :, None], a1[None, :, None]
        b0_, b1_ = b0[:, None, None], b1[:, None, None]
    else:
        a0_, a1_ = a0[None, :], a1[None, :]
        b0_, b1_ = b0[:, None], b1[:, None]
    t = s[:, x0] * a0_ + s[:, x1] * a1_                      # horizontal pass, int
    r0 = (b0_ * (t[y0] >> 4)) >> 16
    r1 = (b1_ * (t[y1] >> 4)) >> 16
    return ((r0 + r1 + 2) >> 2).astype(np.uint8)


# --------------------------------------------------------------------------- a2/a3: histograms


def hist256(plane: np.ndarray) -> np.ndarray:
    """cv2.calcHist([img],[c],None,[256],[0,256]) counts (as int64)."""
    return np.bincount(plane.reshape(-1), minlength=256).astype(np.int64)


def entropy_from_counts(counts: np.ndarray) -> np.float32:
    """-sum p log2 p over p>0, float32 like complexity_metrics.py:412-414."""
    h = counts.astype(np.float32)
    p = h / h.sum()
    p = p[p > 0]
    return np.float32(-np.sum(p * np.log2(p)))


def color_entropy_from_counts(cb, cg, cr):
    """complexity_metrics.py:455-473 (float32, log2(p+1e-8), nan on an empty histogram)."""
    tot = np.float32(0.0)
    for c in (cb, cg, cr):
        h = c.astype(np.float32)
        s = h.sum()
        if s == 0:
            return float("nan")
        p = h / s
        tot = tot + np.sum(p * np.log2(p + 1e-8))
    return np.float32(-tot)


def process_histogram_frame(frame, resize_width, resize_height):
    """complexity_metrics.py:392-416 -- resize THEN gray."""
    g = bgr2gray(resize_linear_u8(frame, resize_width, resize_height))
    return entropy_from_counts(hist256(g))


def process_color_histogram_frame(frame, resize_width, resize_height):
    """complexity_metrics.py:418-475 (CPU branch)."""
    r = resize_linear_u8(frame, resize_width, resize_height)
    return color_entropy_from_counts(hist256(r[..., 0]), hist256(r[..., 1]), hist256(r[..., 2]))


# --------------------------------------------------------------------------- a5/a6: DCT


def dct_matrix(n: int) -> np.ndarray:
    """Orthonormal DCT-II basis, D[k,i] = sqrt(2/n) cos(pi (2i+1) k / 2n), row 0 = sqrt(1/n)
    (what cv2.dct applies along each axis), float64."""
    k = np.arange(n, dtype=np.float64)[:, None]
    i = np.arange(n, dtype=np.float64)[None, :]
    d = np.sqrt(2.0 / n) * np.cos(np.pi * (2.0 * i + 1.0) * k / (2.0 * n))
    d[0, :] = np.sqrt(1.0 / n)
    return d


def dct2(x: np.ndarray) -> np.ndarray:
    """cv2.dct(np.float32(x)) restated as D_h X D_w^T in float64."""
    h, w = x.shape
    return dct_matrix(h) @ x.astype(np.float64) @ dct_matrix(w).T


def dct_input(frame, resize_width, resize_height):
    """gray THEN resize (complexity_metrics.py:358-359, 530-531)."""
    return resize_linear_u8(bgr2gray(frame), resize_width, resize_height)


def process_dct_frame(frame, resize_width, resize_height):
    """complexity_metrics.py:346-364: sum(dct**2).  float64 here; tolerance 1e-4 rel."""
    c = dct2(dct_input(frame, resize_width, resize_height))
    return np.float64(np.sum(c * c))


def process_temporal_dct_frame(prev_gray, curr_gray, resize_width, resize_height):
    """complexity_metrics.py:543-579 (CPU branch): sum |dct(prev) - dct(curr)|."""
    p = resize_linear_u8(prev_gray, resize_width, resize_height)
    c = resize_linear_u8(curr_gray, resize_width, resize_height)
    return np.float64(np.sum(np.abs(dct2(p) - dct2(c))))


# --------------------------------------------------------------------------- a9/a10: series


def process_frame_interval_for_parallel(timestamps):
    """complexity_metrics.py:150-165."""
    prev_t, curr_t = timestamps
    dt = (curr_t - prev_t) / 1000.0
    return 1.0 / dt if dt > 0 else 0.0


def ewm_mean(x, alpha=0.8) -> np.ndarray:
    """pandas Series.ewm(alpha, adjust=True).mean() (complexity_metrics.py:125), float64."""
    x = np.asarray(x, dtype=np.float64)
    out = np.empty_like(x)
    num = den = 0.0
    beta = 1.0 - alpha
    for t, v in enumerate(x):
        num = num * beta + v
        den = den * beta + 1.0
        out[t] = num / den
    return out


def smoothed_mean(x, alpha=0.8):
    """np.mean(smooth_data(x)) (complexity_metrics.py:301-310); nan on an empty series."""
    x = np.asarray(x, dtype=np.float64)
    if x.size == 0:
        return float("nan")
    return float(np.mean(ewm_mean(x, alpha)))


def sampled_indices(n_frames: int, interval: int):
    """read_frame_pairs sampling (complexity_metrics.py:99-107): 0-based I-1, 2I-1, ..."""
    return [(j + 1) * interval - 1 for j in range(n_frames // interval)]


def timestamp_indices(n_frames: int, interval: int):
    """extract_frame_timestamps sampling (complexity_metrics.py:60-69): 0, I, 2I, ..."""
    return list(range(0, n_frames, interval))


# --------------------------------------------------------------------------- a13: PSNR / SSIM


def _plane_sse(a, b) -> int:
    d = a.astype(np.int64) - b.astype(np.int64)
    return int(np.sum(d * d))


def psnr_frame(main_planes, ref_planes):
    """FFmpeg vf_psnr.c (8-bit): per-plane mse, area-weighted mse_avg, 10 log10(255^2/mse).
    Returns dict(mse=[y,u,v], mse_avg, psnr=[y,u,v], psnr_avg)  (SURVEY.md A.9)."""
    areas = [p.shape[0] * p.shape[1] for p in main_planes]
    tot = float(sum(areas))
    mse = [_plane_sse(m, r) / float(a) for m, r, a in zip(main_planes, ref_planes, areas)]
    mse_avg = sum(m * (a / tot) for m, a in zip(mse, areas))

    def ps(m):
        return float("inf") if m == 0 else 10.0 * np.log10(255.0 * 255.0 / m)

    return dict(mse=mse, mse_avg=mse_avg, psnr=[ps(m) for m in mse], psnr_avg=ps(mse_avg))


def ssim_plane(a, b) -> float:
    """FFmpeg vf_ssim.c ssim_plane (8-bit): 4x4 block sums -> overlapping 8x8 windows on a
    4-px grid, ssim_c1 = 416, ssim_c2 = 235963, float32 per-window arithmetic."""
    h, w = a.shape
    bh, bw = h >> 2, w >> 2
    A = a[:bh * 4, :bw * 4].astype(np.int64).reshape(bh, 4, bw, 4)
    B = b[:bh * 4, :bw * 4].astype(np.int64).reshape(bh, 4, bw, 4)
    s1 = A.sum(axis=(1, 3))
    s2 = B.sum(axis=(1, 3))
    ss = (A * A).sum(axis=(1, 3)) + (B * B).sum(axis=(1, 3))
    s12 = (A * B).sum(axis=(1, 3))

    def win(q):
        return q[:-1, :-1] + q[:-1, 1:] + q[1:, :-1] + q[1:, 1:]

    s1, s2, ss, s12 = win(s1), win(s2), win(ss), win(s12)
    vars_ = ss * 64 - s1 * s1 - s2 * s2
    covar = s12 * 64 - s1 * s2
    f = np.float32
    num = (2 * s1 * s2 + 416).astype(f) * (2 * covar + 235963).astype(f)
    den = (s1 * s1 + s2 * s2 + 416).astype(f) * (vars_ + 235963).astype(f)
    v = (num / den).astype(np.float32)
    # float row sums accumulated into a double total (ssim_endn_8bit / ssim_plane)
    rows = np.array([np.float32(np.sum(r.astype(np.float32), dtype=np.float32)) for r in v], dtype=np.float64)
    return float(rows.sum() / ((bh - 1) * (bw - 1)))


def ssim_frame(main_planes, ref_planes):
    areas = [p.shape[0] * p.shape[1] for p in main_planes]
    tot = float(sum(areas))
    s = [ssim_plane(m, r) for m, r in zip(main_planes, ref_planes)]
    return dict(ssim=s, ssim_all=sum(v * (a / tot) for v, a in zip(s, areas)))


def ssim_plane_textbook(a, b) -> float:
    """Independent float64 evaluation of the structural-similarity index (Wang et al. 2004) on vf_ssim.c's
    window set: uniform 8x8 windows on a 4-pixel grid, L = 255, K2 = 0.03 with sample (n-1 = 63) variances.
    Dividing vf_ssim.c's integer expression by 64^2 (mean term) and 64*63 (variance term) gives exactly
        (2 mu_p mu_q + C1) (2 cov + C2) / ((mu_p^2 + mu_q^2 + C1) (var_p + var_q + C2))
    with C2 = (0.03*255)^2 and C1 = (0.01*255)^2 / 64 (ssim_c1 = .01^2 255^2 64 is added to sums that carry a
    factor 64^2 -- x264's constant, kept by FFmpeg), both rounded to integers in the scaled domain.  This
    function shares no code with ssim_plane / the C oracle: it works on pixel windows in floating point, so
    agreement (~1e-6) checks the block sums, the window set and the plane mean; it is a known-answer pin on
    the published definition, not on an FFmpeg binary (none exists in the image or on the GPU box)."""
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    h, w = a.shape
    c1, c2 = 416.0 / 4096.0, 235963.0 / (64.0 * 63.0)
    vals = []
    for y in range(0, (h >> 2) * 4 - 7, 4):
        for x in range(0, (w >> 2) * 4 - 7, 4):
            p, q = a[y:y + 8, x:x + 8], b[y:y + 8, x:x + 8]
            mp, mq = p.mean(), q.mean()
            vp, vq = p.var(ddof=1), q.var(ddof=1)
            cov = ((p - mp) * (q - mq)).sum() / 63.0
            vals.append((2 * mp * mq + c1) * (2 * cov + c2) / ((mp * mp + mq * mq + c1) * (vp + vq + c2)))
    return float(np.mean(vals))


# --------------------------------------------------------------------------- f4: yuv420p -> BGR


def yuv420_to_bgr(y, u, v) -> np.ndarray:
    """cv2.VideoCapture.read() of a yuv420p stream (complexity_metrics.py:38-111 on the ENCODED file,
    video_processing.py:242): libswscale's unscaled yuv420p -> bgr24 converter (x86 SIMD path, taken for every
    even frame size), nearest chroma, 16-bit fixed point with arithmetic shifts:
        yy = ((Y << 3) - 128) * 9539 >> 16 ;  uu = (U << 3) - 1024 ;  vv = (V << 3) - 1024
        B = sat8(yy + (uu * 16525 >> 16)) ; G = sat8(yy + (uu * -3209 >> 16) + (vv * -6660 >> 16)) ;
        R = sat8(yy + (vv * 13075 >> 16))
    Pinned on cv2 itself (yuv4mpeg files through cv2.VideoCapture): all 2^24 (Y,U,V) triples and even sizes
    from 2x2 to 1080p -- oracle/make_golden.py --yuv2bgr, tests/golden/yuv2bgr_cv2.{json,npz}.
    y: [h,w]; u, v: [h/2,w/2] uint8 (or stacks with a leading frame axis).  Returns [...,h,w,3] uint8."""
    y = np.asarray(y)
    h, w = y.shape[-2:]
    if (h | w) & 1:
        raise ValueError("yuv420p -> BGR is defined here for even frame sizes only")
    uf = np.repeat(np.repeat(np.asarray(u), 2, axis=-2), 2, axis=-1).astype(np.int64)
    vf = np.repeat(np.repeat(np.asarray(v), 2, axis=-2), 2, axis=-1).astype(np.int64)
    yy = (((y.astype(np.int64) << 3) - 128) * 9539) >> 16
    uu = (uf << 3) - 1024
    vv = (vf << 3) - 1024
    b = yy + ((uu * 16525) >> 16)
    g = yy + ((uu * -3209) >> 16) + ((vv * -6660) >> 16)
    r = yy + ((vv * 13075) >> 16)
    return np.clip(np.stack([b, g, r], axis=-1), 0, 255).astype(np.uint8)
