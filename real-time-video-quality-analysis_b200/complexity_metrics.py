"""Host mirror of the reference's ``complexity_metrics.py`` -- same function names, argument
meaning, return order/dtypes and error behaviour -- with every pixel operation executed by the
sm_100a kernels of libvqa_b200.so (``_native.py``).  There is no CPU fallback and no
multi-backend switch: the reference's ``use_gpu``/CuPy toggle (complexity_metrics.py:15-22) has
no equivalent here.

Reference interface -> this module (file:line in /root/reference/complexity_metrics.py):
  validate_video_path :25          extract_frame_timestamps :38     read_frame_pairs :76
  smooth_data :114                 process_in_batches :128          process_frame_interval_for_parallel :150
  calculate_scene_complexity_score :171                             calculate_average_scene_complexity :246
  process_frame_complexity :313    process_dct_frame :346           process_orb_frame_for_parallel :367
  process_histogram_frame :392     process_color_histogram_frame :418   process_edge_frame :477
  calculate_temporal_dct :506      process_temporal_dct_frame :543
"""
from __future__ import annotations

import functools
import logging
import multiprocessing

import numpy as np

try:
    from . import _native as N
    from .frame_source import SampledFrameSource
except ImportError:  # used as a top-level drop-in module (package dir on sys.path)
    import _native as N  # type: ignore
    from frame_source import SampledFrameSource  # type: ignore

logging.basicConfig(level=logging.INFO)
logger = logging.getLogger(__name__)

use_gpu = True  # kept for source compatibility; the B200 path is the only path

# ORB input size (width, height).  None = the reference's hard-wired 64x64 (:386), where ORB degenerates
# to a {0, 1} count; a size runs the full pipeline (8-level pyramid, FAST, Harris, retainBest) on
# gray(resize(frame, size)) -- SURVEY.md 8 f2.  Per-call ``orb_size=`` arguments override it.
ORB_SIZE = None


def set_orb_size(size):
    """Module-wide ``orb_size`` knob: (width, height), or None for the reference's 64x64."""
    global ORB_SIZE
    if size is not None:
        size = (int(size[0]), int(size[1]))
        if size[0] <= 0 or size[1] <= 0:
            raise ValueError("orb_size must be positive (width, height)")
    ORB_SIZE = size


def _orb_size(orb_size):
    return ORB_SIZE if orb_size is None else (int(orb_size[0]), int(orb_size[1]))


# --------------------------------------------------------------------------- frame source
def validate_video_path(input_path):
    """Check if the input is a valid video or frame file (reference :25-35)."""
    if not isinstance(input_path, str):
        raise ValueError("Invalid input path. Please provide a valid file path.")
    if input_path.endswith(('.mp4', '.avi', '.mov')):
        return 'video'
    elif input_path.endswith(('.jpg', '.png')):
        return 'frame'
    raise ValueError("Unsupported file type. Please provide a video or frame file.")


def _open_capture(video_path):
    import cv2  # decode stays on OpenCV/libav (frame source is the "next" row)
    return cv2, cv2.VideoCapture(video_path)


def extract_frame_timestamps(video_path, frame_interval=10):
    """Timestamps (ms) of source frames 0, I, 2I, ... (reference :38-73: the modulo test runs
    before the counter is incremented)."""
    validate_video_path(video_path)
    cv2, cap = _open_capture(video_path)
    stamps, count = [], 0
    try:
        if not cap.isOpened():
            logger.error(f"Error opening video file: {video_path}")
            return []
        while cap.isOpened():
            ok, _ = cap.read()
            if not ok:
                break
            if count % frame_interval == 0:
                stamps.append(cap.get(cv2.CAP_PROP_POS_MSEC))
            count += 1
    finally:
        cap.release()
    return stamps


def read_frame_pairs(video_path, frame_interval=10):
    """(current, previous) pairs of the frames at source indices I-1, 2I-1, ... (reference
    :76-111: the counter is incremented before the modulo test)."""
    validate_video_path(video_path)
    cv2, cap = _open_capture(video_path)
    pairs, prev, count = [], None, 0
    try:
        if not cap.isOpened():
            logger.error(f"Error opening video file: {video_path}")
            return []
        while cap.isOpened():
            ok, frame = cap.read()
            if not ok:
                break
            count += 1
            if count % frame_interval == 0:
                if prev is not None:
                    pairs.append((frame, prev))
                prev = frame
    finally:
        cap.release()
    return pairs


# --------------------------------------------------------------------------- series helpers
def smooth_data(data, alpha=0.8):
    """pd.Series(data).ewm(alpha=alpha).mean().to_numpy() (reference :114-125; adjust=True, ignore_na=False),
    float64.  O(T) scalar recurrence on the host; the clip-level mean of the smoothed series
    is reduced on the device (``Context.ewm_partial``).  NaN observations follow pandas: they are left out of
    both sums while the weights of the older terms keep decaying, the previous smoothed value is carried forward,
    and positions before the first valid observation stay NaN."""
    x = np.asarray(list(data) if not isinstance(data, np.ndarray) else data, dtype=np.float64)
    out = np.empty_like(x)
    num = den = 0.0
    beta = 1.0 - alpha
    for t in range(x.shape[0]):
        num *= beta
        den *= beta
        if x[t] == x[t]:
            num += x[t]
            den += 1.0
        out[t] = num / den if den > 0.0 else np.nan
    return out


def _smoothed_mean(series, alpha, empty=float("nan")):
    """np.mean(smooth_data(series, alpha)) evaluated by the device reduction."""
    series = np.asarray(series, dtype=np.float64)
    if series.size == 0:
        return np.float64(empty)
    if np.isnan(series).any():
        # the closed-form weights of the device reduction assume every term is present; with missing observations
        # the reference's value is np.mean of the pandas-smoothed series (NaN iff the series STARTS with NaN)
        return np.float64(np.mean(smooth_data(series, alpha)))
    return np.float64(N.get_context().ewm_partial(series, 0, series.size, alpha))


def process_frame_interval_for_parallel(timestamps):
    """Frame rate between two consecutive timestamps in ms (reference :150-165)."""
    prev_timestamp, curr_timestamp = timestamps
    out = N.get_context().framerate_series([prev_timestamp, curr_timestamp])
    return float(out[0])


# --------------------------------------------------------------------------- per-item operators
def _one(frame, rw, rh, mask, orb_size=None):
    return N.get_context().complexity_frames(np.asarray(frame)[None], rw, rh, mask, orb_size=orb_size)[0]


def process_frame_complexity(frame_pair):
    """Mean Farneback flow magnitude of (current, previous) at native resolution (reference
    :313-343); 0.0 if either frame is None.  np.float32."""
    frame, prev_frame = frame_pair
    if frame is None or prev_frame is None:
        return 0.0
    h, w = frame.shape[:2]
    r = N.get_context().complexity_frames(np.asarray(frame)[None], w, h, N.M_MOTION, halo=np.asarray(prev_frame))
    return np.float32(r["motion"][0])


def process_dct_frame(frame, resize_width, resize_height):
    """sum(dct(resize(gray(frame)))**2) as np.float32 (reference :346-364)."""
    return np.float32(_one(frame, resize_width, resize_height, N.M_DCT)["dct_energy"])


def process_orb_frame_for_parallel(frame, orb_size=None):
    """ORB keypoint count on the 64x64 gray resize (reference :367-389).  int.  ``orb_size`` (or the
    module knob ``ORB_SIZE``) replaces the hard-wired 64x64 by any (width, height)."""
    return int(_one(frame, 64, 64, N.M_ORB, _orb_size(orb_size))["orb_count"])


def process_histogram_frame(frame, resize_width, resize_height):
    """Shannon entropy of the 256-bin gray histogram of the resized frame (reference :392-416)."""
    return np.float32(_one(frame, resize_width, resize_height, N.M_HIST)["hist_entropy"])


def process_color_histogram_frame(frame, resize_width, resize_height):
    """Summed entropies of the B, G, R histograms (reference :418-475); nan on an empty one."""
    return np.float32(_one(frame, resize_width, resize_height, N.M_COLOR)["color_entropy"])


def process_edge_frame(frame, resize_width, resize_height):
    """Number of Canny(100, 200) edge pixels of the resized gray frame (reference :477-504)."""
    return np.int64(_one(frame, resize_width, resize_height, N.M_EDGE)["edge_count"])


def process_temporal_dct_frame(prev_gray_frame, curr_gray_frame, resize_width, resize_height):
    """sum|dct(prev) - dct(curr)| of two GRAY frames (reference :543-579).  The gray planes are
    replicated to B=G=R (gray(v,v,v) == v for every v in the 15-bit formula) so the same ingest
    kernel serves; resize-after-gray order is preserved."""
    p = np.ascontiguousarray(prev_gray_frame, dtype=np.uint8)
    c = np.ascontiguousarray(curr_gray_frame, dtype=np.uint8)
    p3, c3 = np.repeat(p[..., None], 3, axis=2), np.repeat(c[..., None], 3, axis=2)
    r = N.get_context().complexity_frames(c3[None], resize_width, resize_height, N.M_TDCT, halo=p3)
    return np.float32(r["temporal_dct"][0])


# --------------------------------------------------------------------------- batch executor
_FRAME_FIELDS = {
    "process_dct_frame": (N.M_DCT, "dct_energy", np.float32),
    "process_histogram_frame": (N.M_HIST, "hist_entropy", np.float32),
    "process_color_histogram_frame": (N.M_COLOR, "color_entropy", np.float32),
    "process_edge_frame": (N.M_EDGE, "edge_count", np.int64),
    "process_orb_frame_for_parallel": (N.M_ORB, "orb_count", int),
}


def _unwrap(func, kwargs):
    """Peel functools.partial layers; returns (base function, merged keyword arguments)."""
    kw = dict(kwargs)
    while isinstance(func, functools.partial):
        if func.args:
            raise TypeError("process_in_batches: positional partial arguments are not supported")
        kw = {**func.keywords, **kw}
        func = func.func
    return func, kw


def process_in_batches(frames, process_func, num_workers, batch_size=100, **kwargs):
    """Order-preserving batched map (reference :128-148).  The reference spawns a process pool per
    call and pickles each frame to a worker; here ``process_func`` selects a device kernel
    chain and each ``batch_size`` slice is one device batch.  ``num_workers`` is accepted and
    ignored.  Unknown callables raise TypeError (no CPU fallback)."""
    func, kw = _unwrap(process_func, kwargs)
    name = getattr(func, "__name__", repr(func))
    frames = list(frames)
    results = []
    known = set(_FRAME_FIELDS) | {"process_frame_interval_for_parallel", "process_frame_complexity"}
    if name not in known:
        raise TypeError(f"process_in_batches: no device kernel is registered for {name!r}; "
                        "this build has no CPU fallback")
    ctx = N.get_context()
    if name == "process_frame_interval_for_parallel":
        for a, b in frames:
            results.append(float(ctx.framerate_series([a, b])[0]))
        return results
    if name == "process_frame_complexity":
        for i in range(0, len(frames), batch_size):
            batch = frames[i:i + batch_size]
            j = 0
            while j < len(batch):
                cur, prev = batch[j]
                if cur is None or prev is None:
                    results.append(0.0)
                    j += 1
                    continue
                # chain consecutive pairs (pair j's current frame is pair j+1's previous frame)
                k, chain = j, [np.asarray(cur)]
                while k + 1 < len(batch) and batch[k + 1][0] is not None and batch[k + 1][1] is batch[k][0]:
                    chain.append(np.asarray(batch[k + 1][0]))
                    k += 1
                h, w = chain[0].shape[:2]
                r = ctx.complexity_frames(np.stack(chain), w, h, N.M_MOTION, halo=np.asarray(prev))
                results.extend(np.float32(v) for v in r["motion"])
                j = k + 1
        return results
    if name in _FRAME_FIELDS:
        mask, field, cast = _FRAME_FIELDS[name]
        if name == "process_orb_frame_for_parallel":
            rw, rh = 64, 64
        else:
            try:
                rw, rh = int(kw["resize_width"]), int(kw["resize_height"])
            except KeyError as e:
                raise TypeError(f"{name} needs resize_width and resize_height") from e
        for i in range(0, len(frames), batch_size):
            batch = frames[i:i + batch_size]
            shapes = {np.asarray(f).shape for f in batch}
            osz = _orb_size(kw.get("orb_size")) if name == "process_orb_frame_for_parallel" else None
            if len(shapes) == 1:
                r = ctx.complexity_frames(np.stack([np.asarray(f) for f in batch]), rw, rh, mask, orb_size=osz)
                results.extend(cast(v) for v in r[field])
            else:  # ragged batch: one call per frame
                for f in batch:
                    results.append(cast(ctx.complexity_frames(np.asarray(f)[None], rw, rh, mask, orb_size=osz)[field][0]))
        return results
    raise AssertionError("unreachable")


# --------------------------------------------------------------------------- clip level
def _clip_metrics(frames, resize_width, resize_height, batch_size=100, halo=None, mask=N.M_ALL, orb_size=None):
    """All per-frame and pair metrics of consecutive sampled frames in one device pass."""
    ctx = N.get_context()
    return ctx.complexity_frames(frames, resize_width, resize_height, mask, halo=halo, orb_size=_orb_size(orb_size))


def _stack_sampled(frame_pairs):
    """[s_0, s_1, ..., s_{K-1}] from the reference's (s_j, s_{j-1}) pair list."""
    if not frame_pairs:
        return None
    frames = [np.asarray(frame_pairs[0][1])] + [np.asarray(p[0]) for p in frame_pairs]
    return np.stack(frames)


def calculate_temporal_dct(video_path, resize_width, resize_height, frame_interval=10, smoothing_factor=0.8):
    """Average temporal DCT complexity (reference :506-541): consecutive sampled frames of
    pair[0], K-2 values, smoothed mean; 0.0 when empty."""
    pairs = read_frame_pairs(video_path, frame_interval)
    pairs = [p for p in pairs if p[0] is not None and p[1] is not None]
    if len(pairs) < 2:
        return 0.0
    clip = np.stack([np.asarray(p[0]) for p in pairs])
    r = _clip_metrics(clip, resize_width, resize_height, mask=N.M_TDCT)
    return _smoothed_mean(r["temporal_dct"][1:], smoothing_factor, empty=0.0)


STREAM_CHUNK_FRAMES = 48      # sampled frames per decode chunk (= the device chunk of vqa_complexity_frames)


def stream_clip_metrics(video_path, resize_width, resize_height, frame_interval=10, chunk_frames=None, mask=N.M_ALL,
                        orb_size=None):
    """Single-decode streaming analysis of a clip.  Decodes in a background thread,
    pushes chunks of sampled frames through the device with the previous chunk's last frame as
    halo, and returns ``(rows, timestamps)``: one FRAME_DTYPE row per sampled frame ``s_0 .. s_{K-1}``
    (``None`` if the clip yields no sampled frame) and the reference's timestamp list (ms of source
    frames 0, I, 2I, ...).  Results are identical to analysing the fully decoded clip in one call."""
    validate_video_path(video_path)
    src = SampledFrameSource(video_path, frame_interval, chunk_frames or STREAM_CHUNK_FRAMES)
    ctx, parts, halo = None, [], None
    for chunk in src:
        if ctx is None:
            ctx = N.get_context()
        if halo is not None and halo.shape != chunk.shape[1:]:
            # the source flushes a partial chunk when the decoder changes the frame size; the pair across that
            # boundary has no optical flow (cv2.calcOpticalFlowFarneback rejects it in the reference as well)
            raise ValueError("frame size changes mid-stream (%dx%d -> %dx%d) after %d sampled frames: the pair metrics "
                             "are undefined across the change" % (halo.shape[1], halo.shape[0], chunk.shape[2],
                                                                  chunk.shape[1], sum(len(p) for p in parts)))
        parts.append(ctx.complexity_frames(chunk, resize_width, resize_height, mask, halo=halo,
                                           orb_size=_orb_size(orb_size)))
        halo = chunk[-1]
    rows = np.concatenate(parts) if parts else None
    return rows, list(src.timestamps)


def calculate_average_scene_complexity(video_path, resize_width, resize_height, frame_interval=10,
                                       smoothing_factor=0.8, num_workers=None, batch_size=100, *, orb_size=None):
    """Reference :246-310.  Returns, in the reference's order, the means of the EWM-smoothed
    series of: motion, dct, histogram, edge, orb, colour histogram, temporal dct, framerate.
    One decode of the file (the reference makes three) feeds every metric and the timestamps.
    ``orb_size`` (keyword-only extension, default = module knob ``ORB_SIZE`` = the reference's 64x64)."""
    if num_workers is None:
        num_workers = multiprocessing.cpu_count() // 2      # accepted for API compatibility
    a = smoothing_factor
    r, frame_timestamps = stream_clip_metrics(video_path, resize_width, resize_height, frame_interval,
                                              orb_size=orb_size)
    if r is None or len(r) < 2:                              # fewer than one (current, previous) pair
        nan = np.float64("nan")
        motion = dct = hist = edge = orb = color = nan
        tdct = 0.0
    else:
        logger.info("Calculated all scene-complexity metrics on the GPU (%d sampled frames)", len(r))
        # per-frame metrics use pair[0] only: s_1..s_{K-1} (reference :271); s_0 is never analysed
        motion = _smoothed_mean(r["motion"][1:], a)
        dct = _smoothed_mean(r["dct_energy"][1:], a)
        hist = _smoothed_mean(r["hist_entropy"][1:], a)
        edge = _smoothed_mean(r["edge_count"][1:], a)
        orb = _smoothed_mean(r["orb_count"][1:], a)
        color = _smoothed_mean(r["color_entropy"][1:], a)
        tdct = _smoothed_mean(r["temporal_dct"][2:], a, empty=0.0)
    fps = N.get_context().framerate_series(frame_timestamps) if len(frame_timestamps) > 1 else []
    framerate = _smoothed_mean(fps, a)
    return (motion, dct, hist, edge, orb, color, tdct, framerate)


def normalize(value, min_value, max_value):
    """Normalize a value to a 0-1 range based on provided min and max (reference :167-169)."""
    return (value - min_value) / (max_value - min_value) if max_value > min_value else 0


_SCORE_RANGES = (("motion", 0.0, 10.0, 0.25), ("dct", 1e6, 5e7, 0.15), ("hist", 0.0, 8.0, 0.10),
                 ("edge", 0.0, 1.0, 0.10), ("orb", 0.0, 5000, 0.10), ("color", 0.0, 8.0, 0.10),
                 ("tdct", 0.0, 1e7, 0.15), ("fps", 0.0, 2.0, 0.05))


def calculate_scene_complexity_score(encoded_video, resize_width, resize_height, frame_interval=10,
                                     smoothing_factor=0.8, num_workers=None, batch_size=100):
    """Weighted sum of min-max normalised metrics (reference :171-242: ranges :197-206, weights
    :219-228; like the reference it ignores ``num_workers``)."""
    values = calculate_average_scene_complexity(encoded_video, resize_width, resize_height,
                                                frame_interval=frame_interval, smoothing_factor=smoothing_factor,
                                                num_workers=None, batch_size=batch_size)
    return sum(normalize(v, lo, hi) * wgt for v, (_, lo, hi, wgt) in zip(values, _SCORE_RANGES))
