"""Frame-range sharding across the GPUs of one box (one process per GPU) -- SURVEY.md 8(e).

The reference's only parallelism is a per-metric ``ProcessPoolExecutor.map`` over frames
(complexity_metrics.py:143-147).  Here each rank owns a contiguous range of *sampled* frames plus
a one-frame halo (the previous sampled frame) for the two pair metrics, evaluates its rows on its
own GPU with no data-path collective, and the clip-level result -- the mean of the EWM-smoothed
series -- is obtained as a sum of per-rank weighted partial sums (``vqa_ewm_partial``) with ONE
all-reduce of 8 doubles + 3 integers at clip end (NCCL over NVLink on GPUs; gloo in the CPU
tests of this host logic).
"""
from __future__ import annotations

import numpy as np

# series, in the reference's return order (complexity_metrics.py:301-310)
SERIES = ("motion", "dct_energy", "hist_entropy", "edge_count", "orb_count", "color_entropy", "temporal_dct")
# first sampled-frame index that contributes to each series (App. B: s_0 is never analysed,
# the temporal DCT starts one later)
FIRST = {"motion": 1, "dct_energy": 1, "hist_entropy": 1, "edge_count": 1, "orb_count": 1,
         "color_entropy": 1, "temporal_dct": 2}


def shard_range(k_frames: int, rank: int, world: int):
    """Contiguous [a, b) of sampled-frame indices owned by ``rank`` (balanced to +-1)."""
    base, rem = divmod(k_frames, world)
    a = rank * base + min(rank, rem)
    return a, a + base + (1 if rank < rem else 0)


def ewm_coefficients(total: int, alpha: float) -> np.ndarray:
    """c_i with mean(ewm(x)) = sum_i c_i x_i (host closed form; the device kernel evaluates the
    same weights).  Used by the CPU tests of the sharding logic and as a cross-check."""
    if total == 0:
        return np.zeros(0)
    beta = 1.0 - alpha
    t = np.arange(total, dtype=np.float64)
    d = (t + 1.0) if beta == 1.0 else (1.0 - beta ** (t + 1.0)) / (1.0 - beta)
    s = np.zeros(total)
    acc = 0.0
    for i in range(total - 1, -1, -1):
        acc = 1.0 / d[i] + beta * acc
        s[i] = acc
    return s / total


def local_partials(rows, a: int, k_frames: int, alpha: float, partial_fn):
    """Per-series weighted partial sums of this rank's rows.

    rows        structured array for sampled frames a .. a+len(rows)-1 (pair metrics of the
                first row computed against the halo frame when a > 0)
    partial_fn  (x, offset, total, alpha) -> float ; the device reduction in production
    """
    out = np.zeros(len(SERIES), dtype=np.float64)
    for si, name in enumerate(SERIES):
        first = FIRST[name]
        total = max(k_frames - first, 0)
        lo = max(first - a, 0)                      # rows before `first` do not enter the series
        x = np.asarray(rows[name][lo:], dtype=np.float64)
        if total == 0 or x.size == 0:
            continue
        out[si] = partial_fn(x, a + lo - first, total, alpha)
    return out


def init_context_comm(ctx, group=None):
    """Give ``ctx`` its own NCCL communicator over the ranks of the torch process group (vqa_comm_init):
    rank 0 creates the unique id inside the library and the group broadcasts its 128 bytes.  After this,
    ``reduce_partials(..., ctx=ctx)`` runs as ONE ncclAllReduce inside libvqa_b200.so (vqa_clip_reduce) --
    the call a host without any NCCL binding of its own would make."""
    import torch.distributed as dist
    from . import _native as N
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [N.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    ctx.comm_init(rank, world, box[0])
    return ctx


def reduce_partials(partials: np.ndarray, int_totals: np.ndarray, group=None, device=None, ctx=None):
    """Sum the per-rank partial sums (float64) and integer side totals (int64) over all ranks; returns
    numpy arrays.  With a context that owns a communicator (``init_context_comm``) this is vqa_clip_reduce:
    one fused buffer, one ncclAllReduce on the context's stream.  Otherwise (the gloo CPU tests of the
    host logic, or a caller that brought only a torch group) one torch all-reduce of the same fused
    float64 buffer -- integers ride as doubles, exact below 2^53 (checked)."""
    if ctx is not None and getattr(ctx, "comm_world", 0) > 1:
        return ctx.clip_reduce(partials, int_totals)
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return partials, int_totals
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    p = np.asarray(partials, dtype=np.float64)
    i = np.asarray(int_totals, dtype=np.int64)
    if i.size and np.abs(i).max() >= 2 ** 53 // dist.get_world_size(group):
        raise OverflowError("integer total too large for an exact reduce")
    buf = torch.as_tensor(np.concatenate([p.ravel(), i.ravel().astype(np.float64)]), dtype=torch.float64, device=dev)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    out = buf.cpu().numpy()
    return out[:p.size].reshape(p.shape), np.rint(out[p.size:]).astype(np.int64).reshape(i.shape)


def gather_rows(rows, group=None, device=None):
    """Verification alternative of SURVEY.md 8(e): all-gather the per-frame table (K x 7 doubles) so that
    any rank can run the reference's own smoothing on the whole series.  Returns the concatenated
    [K, len(SERIES)] float64 table in rank order (ranks own contiguous, ordered frame ranges)."""
    import torch
    import torch.distributed as dist
    table = np.stack([np.asarray(rows[name], dtype=np.float64) for name in SERIES], axis=1) if len(rows) \
        else np.zeros((0, len(SERIES)))
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return table
    world = dist.get_world_size(group)
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    n = torch.tensor([len(table)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    pad = torch.zeros((max(counts + [1]), len(SERIES)), dtype=torch.float64, device=dev)
    pad[:len(table)] = torch.as_tensor(table, dtype=torch.float64, device=dev)
    parts = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return np.concatenate([p[:c].cpu().numpy() for p, c in zip(parts, counts)], axis=0)


def means_from_table(table: np.ndarray, alpha: float, smooth_fn):
    """The reference's reduction on a gathered table: np.mean(smooth_data(series)) per column
    (complexity_metrics.py:301-310) with the series starts of App. B; ``smooth_fn(x, alpha)`` -> array."""
    out = []
    for si, name in enumerate(SERIES):
        x = table[FIRST[name]:, si]
        out.append(float(np.mean(smooth_fn(x, alpha))) if len(x) else (0.0 if name == "temporal_dct" else float("nan")))
    return out


def finalize(partials: np.ndarray, k_frames: int, framerate_mean: float):
    """8-tuple in the reference's order; empty series -> nan (temporal DCT -> 0.0)."""
    vals = []
    for si, name in enumerate(SERIES):
        total = k_frames - FIRST[name]
        if total <= 0:
            vals.append(0.0 if name == "temporal_dct" else float("nan"))
        else:
            vals.append(float(partials[si]))
    return tuple(np.float64(v) for v in vals) + (np.float64(framerate_mean),)


def sharded_average_scene_complexity(local_frames, a: int, k_frames: int, resize_width: int, resize_height: int,
                                     timestamps_ms, halo=None, alpha: float = 0.8, group=None, ctx=None):
    """Multi-GPU ``calculate_average_scene_complexity`` for pre-decoded sampled frames.

    local_frames  (m,h,w,3) uint8 (host or CUDA tensor): sampled frames a .. a+m-1 of the clip
    halo          sampled frame a-1 (required when a > 0)
    timestamps_ms the clip's sampled timestamps (every rank holds them: a few bytes per frame)
    """
    from . import _native as N
    ctx = ctx or N.get_context()
    rows = ctx.complexity_frames(local_frames, resize_width, resize_height, N.M_ALL, halo=halo if a > 0 else None)
    partials = local_partials(rows, a, k_frames, alpha, ctx.ewm_partial)
    ints = np.array([int(rows["edge_count"][max(1 - a, 0):].sum()), int(rows["orb_count"][max(1 - a, 0):].sum()),
                     len(rows)], dtype=np.int64)
    partials, ints = reduce_partials(partials, ints, group, ctx=ctx)
    fps = ctx.framerate_series(timestamps_ms) if len(timestamps_ms) > 1 else np.zeros(0)
    fr = ctx.ewm_partial(fps, 0, len(fps), alpha) if len(fps) else float("nan")
    return finalize(partials, k_frames, fr), ints


# --------------------------------------------------------------------------- many clips (BASELINE config 5)
def plan_clip_shards(clip_frames, world: int):
    """"shard by clip first, then by frame range inside a clip if clips < GPUs".

    clip_frames  sampled-frame count K_c of every clip
    returns      per rank, a list of (clip, a, b): sampled frames [a, b) of that clip.

    clips >= ranks: whole clips, longest first onto the least loaded rank (no halo anywhere).
    clips <  ranks: every clip gets a group of ranks in proportion to its length (at least one) and is
    cut into contiguous ranges inside the group (one halo frame per cut).  Deterministic, so every
    rank computes the same plan without communication."""
    n = len(clip_frames)
    plan = [[] for _ in range(world)]
    if n == 0:
        return plan
    if n >= world:
        load = [0] * world
        for c in sorted(range(n), key=lambda i: (-clip_frames[i], i)):
            if clip_frames[c] <= 0:                       # an empty clip has nothing to shard
                continue
            r = min(range(world), key=lambda j: (load[j], j))
            plan[r].append((c, 0, int(clip_frames[c])))
            load[r] += int(clip_frames[c])
        for p in plan:
            p.sort()
        return plan
    total = float(sum(clip_frames)) or 1.0
    ranks = [1] * n
    for _ in range(world - n):                       # hand out the spare ranks to the clip with the most frames per rank
        c = max(range(n), key=lambda i: (clip_frames[i] / ranks[i], -i))
        ranks[c] += 1
    r0 = 0
    for c in range(n):
        for j in range(ranks[c]):
            a, b = shard_range(int(clip_frames[c]), j, ranks[c])
            if b > a:
                plan[r0 + j].append((c, a, b))
        r0 += ranks[c]
    return plan


def multi_clip_partials(plan_for_rank, rows_of, clip_frames, alpha: float, partial_fn):
    """[n_clips, 7] weighted partial sums and [n_clips, 3] integer totals (edges, ORB keypoints, frames)
    of this rank's shards.  ``rows_of(clip, a, b)`` returns the FRAME_DTYPE rows of sampled frames
    [a, b) of ``clip`` (pair metrics of row 0 against frame a-1 when a > 0)."""
    n = len(clip_frames)
    partials = np.zeros((n, len(SERIES)), dtype=np.float64)
    ints = np.zeros((n, 3), dtype=np.int64)
    for clip, a, b in plan_for_rank:
        rows = rows_of(clip, a, b)
        partials[clip] += local_partials(rows, a, int(clip_frames[clip]), alpha, partial_fn)
        lo = max(1 - a, 0)
        ints[clip] += (int(rows["edge_count"][lo:].sum()), int(rows["orb_count"][lo:].sum()), len(rows))
    return partials, ints


def sharded_multi_clip_scene_complexity(clips, resize_width: int, resize_height: int, timestamps_ms, rank: int, world: int,
                                        alpha: float = 0.8, group=None, ctx=None, clip_frames=None):
    """All clips of a batch over all ranks: ONE all-reduce of [n_clips x 7] doubles (+ [n_clips x 3]
    integers) closes every clip.  ``clips[c]`` is the (K_c,h,w,3) uint8 array (host or CUDA tensor) of
    sampled frames -- only the shards of ``plan_clip_shards(...)[rank]`` are touched, so a rank may
    pass ``None`` for clips it does not own.  ``clip_frames[c]`` = K_c, the number of SAMPLED FRAMES of clip c
    (floor(N/I), App. B) -- required whenever a rank holds ``None`` for a clip, and NOT the same as
    ``len(timestamps_ms[c])`` (ceil(N/I) entries: one more than frames when N % I != 0).  Returns a list of
    8-tuples in the reference's order."""
    from . import _native as N
    ctx = ctx or N.get_context()
    if clip_frames is None:
        if any(c is None for c in clips):
            raise ValueError("clip_frames is required when some clips are not held by this rank")
        clip_frames = [len(c) for c in clips]
    clip_frames = [int(k) for k in clip_frames]
    if len(clip_frames) != len(clips) or len(timestamps_ms) != len(clips):
        raise ValueError("clips, clip_frames and timestamps_ms must have one entry per clip")
    for c, fr in enumerate(clips):
        if fr is not None and len(fr) != clip_frames[c]:
            raise ValueError(f"clip {c}: {len(fr)} frames held but clip_frames says {clip_frames[c]}")
    plan = plan_clip_shards(clip_frames, world)[rank]

    def rows_of(clip, a, b):
        fr = clips[clip]
        if fr is None:
            raise ValueError(f"rank {rank} owns frames [{a},{b}) of clip {clip} but was not given the clip")
        return ctx.complexity_frames(fr[a:b], resize_width, resize_height, N.M_ALL, halo=fr[a - 1] if a > 0 else None)

    partials, ints = multi_clip_partials(plan, rows_of, clip_frames, alpha, ctx.ewm_partial)
    partials, ints = reduce_partials(partials, ints, group, ctx=ctx)
    out = []
    for c, ts in enumerate(timestamps_ms):
        fps = ctx.framerate_series(ts) if len(ts) > 1 else np.zeros(0)
        fr = ctx.ewm_partial(fps, 0, len(fps), alpha) if len(fps) else float("nan")
        out.append(finalize(partials[c], clip_frames[c], fr))
    return out, ints
