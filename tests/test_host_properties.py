"""Property tests (hypothesis) of the host-side arithmetic that multi-GPU correctness rests on: the closed-form
EWM-mean coefficients, contiguous frame ranges and the clip-first shard plan (8e).  CPU only."""
import numpy as np
import pytest

hyp = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st   # noqa: E402

from oracle import np_oracle as NO   # noqa: E402


@settings(max_examples=60, deadline=None)
@given(st.lists(st.floats(-1e6, 1e6, allow_nan=False, width=32), min_size=1, max_size=200),
       st.sampled_from([0.8, 0.3, 0.05, 1.0]), st.integers(1, 9))
def test_ewm_coefficients_are_the_smoothed_mean_for_any_split(vqa, xs, alpha, world):
    """sum_i c_i x_i == mean(ewm(x)) (pandas adjust=True closed form) and is invariant to the shard count."""
    from rtvqa_b200 import sharding as SH
    x = np.asarray(xs, dtype=np.float64)
    c = SH.ewm_coefficients(len(x), alpha)
    assert c.sum() == pytest.approx(1.0, rel=1e-12)            # a weighted mean
    whole = float(np.dot(c, x))
    scale = max(1.0, float(np.abs(x).max()))
    assert whole == pytest.approx(NO.smoothed_mean(x, alpha), abs=1e-9 * scale)
    parts = 0.0
    for r in range(world):
        a, b = SH.shard_range(len(x), r, world)
        parts += float(np.dot(c[a:b], x[a:b]))
    assert parts == pytest.approx(whole, abs=1e-9 * scale)


@settings(max_examples=100, deadline=None)
@given(st.integers(0, 5000), st.integers(1, 64))
def test_shard_ranges_partition_the_clip(vqa, k, world):
    from rtvqa_b200 import sharding as SH
    ranges = [SH.shard_range(k, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == k
    assert all(a <= b for a, b in ranges) and all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    sizes = [b - a for a, b in ranges]
    assert max(sizes) - min(sizes) <= 1


@settings(max_examples=100, deadline=None)
@given(st.lists(st.integers(0, 2000), min_size=0, max_size=40), st.integers(1, 16))
def test_clip_shard_plan_covers_every_frame_exactly_once(vqa, clip_frames, world):
    from rtvqa_b200 import sharding as SH
    plan = SH.plan_clip_shards(clip_frames, world)
    assert len(plan) == world and plan == SH.plan_clip_shards(list(clip_frames), world)
    seen = {c: [] for c in range(len(clip_frames))}
    for shards in plan:
        for c, a, b in shards:
            assert 0 <= a < b <= clip_frames[c]
            seen[c].append((a, b))
    for c, k in enumerate(clip_frames):
        rs = sorted(seen[c])
        assert sum(b - a for a, b in rs) == k
        assert all(rs[i][1] == rs[i + 1][0] for i in range(len(rs) - 1))
    if len(clip_frames) >= world and clip_frames:
        load = [sum(b - a for _, a, b in shards) for shards in plan]
        assert max(load) - min(load) <= max(clip_frames)       # greedy longest-first bound
        assert all(a == 0 and b == clip_frames[c] for shards in plan for c, a, b in shards)


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 4096), st.integers(1, 4096))
def test_exact_taps_stay_inside_the_source(vqa, sn, dn):
    """INTER_LINEAR_EXACT taps of the ABI: offsets inside the source, 8.8 weights in [0, 256], edge replication."""
    from rtvqa_b200 import _native as N
    off, c1 = N.exact_taps(sn, dn)
    assert off.min() >= 0 and off.max() <= sn - 1 and c1.min() >= 0 and c1.max() <= 256
    assert np.all(np.diff(off) >= 0)                           # monotone sampling positions
    assert np.all(c1[off == sn - 1] == 0)                      # the last source pixel is only ever replicated
