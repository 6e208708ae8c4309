"""General-size ORB on the device (csrc/orb.cu) against the oracle and against fixtures
made by cv2's ORB (tests/golden/orb_general.json).  Everything here is integer / bit-pattern work:
pyramid levels, keypoint sets, FAST scores, Harris responses and counts must be identical."""
import numpy as np
import pytest

from oracle import np_oracle as NO
from oracle import orb_oracle as OO
from test_orb_cpu import CFGS, digest, digest_keypoints, golden_gray, orb_golden  # noqa: F401  (fixture re-export)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(vqa):
    from rtvqa_b200 import _native as N
    c = N.Context(0)
    yield c
    c.close()


def _cfg(N, kw):
    return N.orb_cfg(kw.get("nfeatures", 500), kw.get("scale_factor", 1.2), kw.get("nlevels", 8),
                     kw.get("edge_threshold", 31), kw.get("fast_threshold", 20))


def _rows(kps):
    """(octave, x, y, response bits, fast score) set of a KEYPOINT_DTYPE array."""
    return sorted((int(k["octave"]), int(k["lx"]), int(k["ly"]), int(np.float32(k["response"]).view(np.uint32)),
                   int(k["fast_score"])) for k in kps)


def _oracle_rows(rows):
    return sorted((l, x, y, int(np.float32(r).view(np.uint32)), s) for l, x, y, r, s in rows)


@pytest.mark.parametrize("h,w", [(96, 128), (270, 480), (201, 333), (75, 101)])
def test_pyramid_levels_bit_exact(ctx, h, w):
    g = np.random.default_rng(h * 7 + w).integers(0, 256, (h, w), dtype=np.uint8)
    want = OO.pyramid(g)
    for level in range(1, 8):
        if min(want[level].shape) <= 62:          # levels that cannot hold a keypoint are never built
            break
        got = ctx.debug_orb_pyramid(g, level)
        assert got.shape == want[level].shape
        assert np.array_equal(got, want[level]), f"level {level} of {h}x{w}"


def test_keypoints_match_cv2_fixtures(ctx, orb_golden, small_clip, synth):
    from rtvqa_b200 import _native as N
    for case in orb_golden["cases"]:
        gray = golden_gray(case, small_clip, synth)
        kw = CFGS[case["cfg"]]
        counts, levels, kps = ctx.orb_detect(gray, _cfg(N, kw), keypoints=True)
        tag = (case["clip"], case["frame"], case["cfg"])
        assert int(counts[0]) == case["count"], tag
        assert levels[0, :len(case["per_level"])].tolist() == case["per_level"], tag
        assert digest([(int(k["octave"]), 0, 0, float(k["response"])) for k in kps[0]]) == case["digest"], tag
        # every cv2.KeyPoint field (pt, size, angle, response, octave), float32 bit patterns
        full = [(int(k["octave"]), float(k["x"]), float(k["y"]), float(k["size"]), float(k["angle"]), float(k["response"]))
                for k in kps[0]]
        assert digest_keypoints(full) == case["digest_keypoints"], tag


@pytest.mark.parametrize("h,w,kind", [(96, 128, "clip"), (270, 480, "clip"), (200, 333, "noise"), (150, 150, "blur"),
                                      (300, 64, "noise"), (63, 400, "noise")])
def test_keypoint_sets_match_oracle(ctx, synth, h, w, kind):
    """Every keypoint: level, integer position, FAST score and the Harris response bit pattern."""
    rng = np.random.default_rng(h + 3 * w)
    if kind == "clip":
        g = NO.bgr2gray(synth.synth_clip(2, h, w, seed=h)[1])
    else:
        g = rng.integers(0, 256, (h, w), dtype=np.uint8)
        if kind == "blur":
            k = np.array([1, 4, 6, 4, 1], np.int64)
            t = np.apply_along_axis(lambda r: np.convolve(r, k, "same"), 1, g.astype(np.int64))
            g = (np.apply_along_axis(lambda c: np.convolve(c, k, "same"), 0, t) // 256).astype(np.uint8)
    counts, levels, kps = ctx.orb_detect(g, keypoints=True)
    rows, per = OO.orb_detect(g)
    assert levels[0, :8].tolist() == per and int(counts[0]) == sum(per)
    assert _rows(kps[0]) == _oracle_rows(rows)
    want = sorted(OO.orb_keypoints(g))
    got = sorted((int(k["octave"]), float(k["x"]), float(k["y"]), float(k["size"]), float(k["angle"]), float(k["response"]))
                 for k in kps[0])
    assert got == want                                                  # pt, size, orientation, response: exact


def test_batches_device_input_and_flat_frames(ctx, synth):
    """n > the internal batch of 16, device-resident input, frames with zero keypoints in between."""
    import torch
    clip = synth.synth_clip(6, 144, 192, seed=9)
    grays = [NO.bgr2gray(f) for f in clip]
    grays.insert(2, np.full((144, 192), 90, np.uint8))
    stack = np.stack(grays * 3)                                    # 21 frames
    want = [sum(OO.orb_detect(g)[1]) for g in grays] * 3
    counts, _ = ctx.orb_detect(stack)
    assert counts.tolist() == want and want[2] == 0
    counts_dev, _ = ctx.orb_detect(torch.from_numpy(stack).cuda())
    assert counts_dev.tolist() == want
    from rtvqa_b200 import _native as N
    with pytest.raises(N.VqaError):
        ctx.orb_detect(stack, N.orb_cfg(edge_threshold=2))
    with pytest.raises(N.VqaError):
        ctx.orb_detect(np.zeros((1, 8, 5000), np.uint8))


def test_orb_size_knob_through_complexity_frames(ctx, synth, small_clip):
    """vqa_complexity_frames with orb_width/orb_height: gray(resize(frame, size)) -> full ORB, next to the
    other metrics, for native size, a down-scale, the explicit 64x64 and the default."""
    from rtvqa_b200 import _native as N
    clip = synth.synth_clip(4, 270, 480, seed=3)
    base = ctx.complexity_frames(clip, 64, 64)
    native = ctx.complexity_frames(clip, 64, 64, orb_size=(480, 270))
    small = ctx.complexity_frames(clip, 64, 64, orb_size=(320, 200))
    only = ctx.complexity_frames(clip, 64, 64, N.M_ORB, orb_size=(480, 270))
    same = ctx.complexity_frames(clip, 64, 64, orb_size=(64, 64))
    for i, f in enumerate(clip):
        assert int(native["orb_count"][i]) == OO.orb_count(NO.bgr2gray(f))
        assert int(small["orb_count"][i]) == OO.orb_count(NO.bgr2gray(NO.resize_linear_u8(f, 320, 200)))
        assert int(only["orb_count"][i]) == int(native["orb_count"][i])
        assert int(same["orb_count"][i]) == int(base["orb_count"][i])
    for name in ("edge_count", "hist_entropy", "dct_energy", "motion"):           # the knob touches nothing else
        assert np.array_equal(native[name], base[name], equal_nan=True)
    # identity resize path (rw, rh = native) + native ORB share the full-resolution gray plane
    ident = ctx.complexity_frames(clip, 480, 270, orb_size=(480, 270))
    assert ident["orb_count"].tolist() == native["orb_count"].tolist()


def test_hd_and_4k_counts(ctx, synth):
    hd = synth.synth_clip(2, 1080, 1920, seed=0)
    rows = ctx.complexity_frames(hd, 64, 64, orb_size=(1920, 1080))
    g = NO.bgr2gray(hd[1])
    rows_o, per = OO.orb_detect(g)
    assert int(rows["orb_count"][1]) == sum(per)
    counts, levels, kps = ctx.orb_detect(g, keypoints=True)
    assert _rows(kps[0]) == _oracle_rows(rows_o)
    uhd = NO.bgr2gray(synth.synth_clip(1, 2160, 3840, seed=2)[0])
    counts, levels = ctx.orb_detect(uhd)
    assert levels[0, :8].tolist() == OO.orb_detect(uhd)[1]


def test_drop_in_orb_size(vqa, ctx, synth):
    from rtvqa_b200 import complexity_metrics as cm
    clip = synth.synth_clip(3, 144, 192, seed=4)
    want = [OO.orb_count(NO.bgr2gray(f)) for f in clip]
    assert [cm.process_orb_frame_for_parallel(f, orb_size=(192, 144)) for f in clip] == want
    got = cm.process_in_batches(list(clip), cm.process_orb_frame_for_parallel, 2, orb_size=(192, 144))
    assert got == want and all(isinstance(v, int) for v in got)
    cm.set_orb_size((192, 144))
    try:
        assert cm.process_in_batches(list(clip), cm.process_orb_frame_for_parallel, 2) == want
    finally:
        cm.set_orb_size(None)
    assert cm.process_orb_frame_for_parallel(clip[0]) in (0, 1)                  # reference behaviour restored
