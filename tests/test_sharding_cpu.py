"""Frame-range sharding host logic on CPU: world_size-2 gloo run of the partial-sum reduce
 with the oracle standing in for the per-frame kernels, plus shard-count invariance
of the closed-form EWM coefficients."""
import os
import socket

import numpy as np
import pytest

from oracle import np_oracle as NO
from oracle import ref_port as RP


def _rows_from_oracle(clip, a, halo, rw, rh, dtype):
    rows = np.zeros(len(clip), dtype=dtype)
    prev = halo
    for i, f in enumerate(clip):
        rows["hist_entropy"][i] = RP.o_hist(f, rw, rh)
        rows["color_entropy"][i] = RP.o_color(f, rw, rh)
        rows["dct_energy"][i] = RP.o_dct(f, rw, rh)
        rows["edge_count"][i] = RP.o_edge(f, rw, rh)
        rows["orb_count"][i] = RP.o_orb(f)
        if prev is not None:
            rows["motion"][i] = RP.o_motion((f, prev))
            rows["temporal_dct"][i] = RP.o_tdct(NO.dct_input(prev, rw, rh), NO.dct_input(f, rw, rh), rw, rh)
        else:
            rows["motion"][i] = rows["temporal_dct"][i] = np.nan
        prev = f
    return rows


def _host_partial(x, offset, total, alpha):
    import rtvqa_b200
    from rtvqa_b200 import sharding as SH
    c = SH.ewm_coefficients(total, alpha)
    return float(np.dot(c[offset:offset + len(x)], x))


def _worker(rank, world, port, clip, out_path):
    import torch.distributed as dist
    import rtvqa_b200
    from rtvqa_b200 import _native as N
    from rtvqa_b200 import sharding as SH
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    k = len(clip)
    a, b = SH.shard_range(k, rank, world)
    halo = clip[a - 1] if a > 0 else None
    rows = _rows_from_oracle(clip[a:b], a, halo, 64, 64, N.FRAME_DTYPE)
    partials = SH.local_partials(rows, a, k, 0.8, _host_partial)
    ints = np.array([int(rows["edge_count"][max(1 - a, 0):].sum()), int(rows["orb_count"][max(1 - a, 0):].sum()), len(rows)], np.int64)
    partials, ints = SH.reduce_partials(partials, ints)
    ts = 1000.0 * np.arange(k) / 30.0
    fps = [NO.process_frame_interval_for_parallel((x, y)) for x, y in zip(ts[:-1], ts[1:])]
    res = SH.finalize(partials, k, NO.smoothed_mean(fps, 0.8))
    # verification alternative: all-gather the per-frame table, smooth the whole series
    table = SH.gather_rows(rows)
    gathered = SH.means_from_table(table, 0.8, NO.ewm_mean)
    if rank == 0:
        np.save(out_path, np.array(list(res) + [float(v) for v in ints] + gathered + [float(len(table))]))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_gloo_reduce_matches_single_pass(vqa, small_clip, golden, tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.npy")
    clip = small_clip[:8]
    mp.spawn(_worker, args=(2, _free_port(), clip, out), nprocs=2, join=True)
    got = np.load(out)
    # (1) shard-count invariance: the single-rank evaluation of the same rows, <= 1e-12 relative
    from rtvqa_b200 import _native as N
    from rtvqa_b200 import sharding as SH
    rows = _rows_from_oracle(clip, 0, None, 64, 64, N.FRAME_DTYPE)
    one = SH.finalize(SH.local_partials(rows, 0, len(clip), 0.8, _host_partial), len(clip), got[7])
    np.testing.assert_allclose(got[:8], one, rtol=1e-12)
    # (2) and it is the reference's number (rows carry float32 fields like the reference's np.float32 returns)
    want = RP.average_scene_complexity(clip, 64, 64, frame_interval=1)
    np.testing.assert_allclose(got[:8], want, rtol=1e-6)
    edges = [RP.o_edge(f, 64, 64) for f in clip[1:]]
    assert got[8] == sum(edges) and got[10] == len(clip)          # integer totals identical for any rank count
    # (3) the all-gather + whole-series smoothing path agrees with the partial-sum all-reduce to <= 1e-12
    assert got[18] == len(clip)
    np.testing.assert_allclose(got[11:18], got[:7], rtol=1e-12)


@pytest.mark.parametrize("k,world", [(30, 1), (30, 2), (30, 4), (30, 8), (7, 8), (299, 8)])
def test_shard_ranges_and_partial_sum_invariance(vqa, k, world):
    from rtvqa_b200 import sharding as SH
    ranges = [SH.shard_range(k, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == k
    assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    assert max(b - a for a, b in ranges) - min(b - a for a, b in ranges) <= 1
    x = np.random.default_rng(k).normal(size=k)
    for alpha in (0.8, 0.3, 1.0):
        c = SH.ewm_coefficients(k, alpha)
        whole = float(np.dot(c, x))
        assert whole == pytest.approx(NO.smoothed_mean(x, alpha), rel=1e-13)
        parts = sum(float(np.dot(c[a:b], x[a:b])) for a, b in ranges)
        assert parts == pytest.approx(whole, rel=1e-13)


def test_empty_series_conventions(vqa):
    from rtvqa_b200 import sharding as SH
    res = SH.finalize(np.zeros(len(SH.SERIES)), 1, float("nan"))      # one sampled frame: nothing analysed
    assert all(np.isnan(v) for v in res[:6]) and res[6] == 0.0 and np.isnan(res[7])
    res = SH.finalize(np.ones(len(SH.SERIES)), 2, 3.0)                 # two frames: temporal DCT still empty -> 0.0
    assert res[0] == 1.0 and res[6] == 0.0 and res[7] == 3.0


# ------------------------------------------------------------------ many clips (BASELINE config 5)
@pytest.mark.parametrize("clip_frames,world", [([600] * 64, 8), ([600, 300, 900, 120, 60], 2), ([30, 10], 8),
                                               ([299], 8), ([5, 0, 7], 4), ([], 4), ([3], 8)])
def test_clip_shard_plan_covers_every_frame_once(vqa, clip_frames, world):
    from rtvqa_b200 import sharding as SH
    plan = SH.plan_clip_shards(clip_frames, world)
    assert len(plan) == world
    seen = {c: [] for c in range(len(clip_frames))}
    for shards in plan:
        for c, a, b in shards:
            assert 0 <= a < b <= clip_frames[c]
            seen[c].append((a, b))
    for c, k in enumerate(clip_frames):
        rs = sorted(seen[c])
        assert sum(b - a for a, b in rs) == k
        assert all(rs[i][1] == rs[i + 1][0] for i in range(len(rs) - 1)) and (not rs or (rs[0][0] == 0 and rs[-1][1] == k))
    if len(clip_frames) >= world:                                  # whole clips only: no halo needed anywhere
        assert all(a == 0 and b == clip_frames[c] for shards in plan for c, a, b in shards)
        load = [sum(b - a for _, a, b in shards) for shards in plan]
        assert max(load) - min(load) <= max(clip_frames)
    assert plan == SH.plan_clip_shards(clip_frames, world)         # deterministic: every rank derives the same plan


def test_multi_clip_partials_equal_per_clip_single_pass(vqa):
    """Any rank count gives the same per-clip means as one pass per clip (<= 1e-12), integers identical."""
    from rtvqa_b200 import _native as N
    from rtvqa_b200 import sharding as SH
    rng = np.random.default_rng(5)
    lens = [17, 4, 9]
    tables = []
    for k in lens:
        t = np.zeros(k, dtype=N.FRAME_DTYPE)
        for name in SH.SERIES:
            t[name] = rng.integers(0, 1000, k) if name in ("edge_count", "orb_count") else rng.normal(size=k) * 10
        tables.append(t)

    def rows_of(clip, a, b):
        return tables[clip][a:b]

    want_p, want_i = SH.multi_clip_partials(SH.plan_clip_shards(lens, 1)[0], rows_of, lens, 0.8, _host_partial)
    for world in (2, 3, 8):
        plan = SH.plan_clip_shards(lens, world)
        parts = [SH.multi_clip_partials(plan[r], rows_of, lens, 0.8, _host_partial) for r in range(world)]
        got_p, got_i = sum(p for p, _ in parts), sum(i for _, i in parts)
        np.testing.assert_allclose(got_p, want_p, rtol=1e-12)
        assert np.array_equal(got_i, want_i)
    for c, k in enumerate(lens):                                   # and it is the reference's smoothed mean
        for si, name in enumerate(SH.SERIES):
            x = np.asarray(tables[c][name][SH.FIRST[name]:], dtype=np.float64)
            assert want_p[c, si] == pytest.approx(NO.smoothed_mean(x, 0.8), rel=1e-12)


def _multi_worker(rank, world, port, clips, out_path):
    import torch.distributed as dist
    import rtvqa_b200
    from rtvqa_b200 import _native as N
    from rtvqa_b200 import sharding as SH
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lens = [len(c) for c in clips]
    plan = SH.plan_clip_shards(lens, world)[rank]

    def rows_of(clip, a, b):
        return _rows_from_oracle(clips[clip][a:b], a, clips[clip][a - 1] if a > 0 else None, 64, 64, N.FRAME_DTYPE)

    partials, ints = SH.multi_clip_partials(plan, rows_of, lens, 0.8, _host_partial)
    partials, ints = SH.reduce_partials(partials, ints)
    if rank == 0:
        np.save(out_path, np.concatenate([partials.ravel(), ints.ravel().astype(np.float64)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cuts", [((0, 5), (5, 8), (8, 12)), ((0, 9),)])
def test_two_rank_gloo_multi_clip_reduce(vqa, small_clip, tmp_path, cuts):
    """BASELINE config 5 over 2 gloo ranks: three clips placed whole (clips >= ranks) and one clip cut
    into frame ranges with a halo (clips < ranks); ONE all-reduce of [n_clips x 7] closes the batch and
    every clip's means equal the reference-shaped single pass."""
    import torch.multiprocessing as mp
    from rtvqa_b200 import sharding as SH
    clips = [small_clip[a:b] for a, b in cuts]
    out = str(tmp_path / "multi.npy")
    mp.spawn(_multi_worker, args=(2, _free_port(), clips, out), nprocs=2, join=True)
    got = np.load(out)
    n = len(clips)
    partials, ints = got[:n * len(SH.SERIES)].reshape(n, -1), got[n * len(SH.SERIES):].reshape(n, 3)
    for c, clip in enumerate(clips):
        want = RP.average_scene_complexity(clip, 64, 64, frame_interval=1)
        res = SH.finalize(partials[c], len(clip), want[7])
        np.testing.assert_allclose(res, want, rtol=1e-6)
        assert ints[c, 0] == sum(RP.o_edge(f, 64, 64) for f in clip[1:]) and ints[c, 2] == len(clip)


def test_multi_clip_needs_frame_counts_not_timestamp_counts(vqa):
    """App. B: a clip of N source frames sampled every I has floor(N/I) sampled frames but ceil(N/I) timestamps, so
    with N % I != 0 there is one more timestamp than frames.  The clip plan must be built from the FRAME count
    (ADVICE r1): `clip_frames` is explicit, inconsistent inputs are refused before any device work."""
    from rtvqa_b200 import sharding as SH
    frames = np.zeros((3, 8, 8, 3), np.uint8)                 # N = 35, I = 10 -> 3 sampled frames, 4 timestamps
    ts = [0.0, 333.3, 666.7, 1000.0]
    with pytest.raises(ValueError, match="clip_frames is required"):
        SH.sharded_multi_clip_scene_complexity([None], 64, 64, [ts], rank=0, world=1, ctx=object())
    with pytest.raises(ValueError, match="frames held but clip_frames says"):
        SH.sharded_multi_clip_scene_complexity([frames], 64, 64, [ts], rank=0, world=1, ctx=object(), clip_frames=[len(ts)])
    with pytest.raises(ValueError, match="one entry per clip"):
        SH.sharded_multi_clip_scene_complexity([frames], 64, 64, [ts, ts], rank=0, world=1, ctx=object(), clip_frames=[3])
    # the plan follows the frame count: 3 frames over 4 ranks leave the last rank empty, never a phantom 4th frame
    plan = SH.plan_clip_shards([3], 4)
    assert [p for p in plan if p] == [[(0, 0, 1)], [(0, 1, 2)], [(0, 2, 3)]]


def _fused_worker(rank, world, port, out_path):
    import torch.distributed as dist
    import rtvqa_b200  # noqa: F401
    from rtvqa_b200 import sharding as SH
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = np.arange(14, dtype=np.float64).reshape(2, 7) * (rank + 1) + 0.125
    i = np.array([[2 ** 40 + rank, 7, 3], [5, rank, 1]], dtype=np.int64)
    gp, gi = SH.reduce_partials(p, i)
    try:
        SH.reduce_partials(p, np.array([2 ** 53], dtype=np.int64))
        overflow = False
    except OverflowError:
        overflow = True
    if rank == 0:
        np.savez(out_path, p=gp, i=gi, overflow=overflow)
    dist.barrier()
    dist.destroy_process_group()


def test_fused_reduce_keeps_shapes_and_exact_integers(vqa, tmp_path):
    """reduce_partials moves ONE fused float64 buffer (doubles + integers as doubles): shapes survive, integers
    stay exact, and an integer that could round is refused instead of rounded."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "fused.npz")
    mp.spawn(_fused_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    z = np.load(out)
    assert z["p"].shape == (2, 7) and z["i"].shape == (2, 3) and z["i"].dtype == np.int64
    np.testing.assert_array_equal(z["p"], np.arange(14, dtype=np.float64).reshape(2, 7) * 3 + 0.25)
    assert z["i"].tolist() == [[2 ** 41 + 1, 14, 6], [10, 1, 2]] and bool(z["overflow"])
