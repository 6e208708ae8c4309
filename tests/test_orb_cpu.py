"""General-size ORB, CPU side: the oracle (oracle/orb_oracle.py) against fixtures made
by cv2's ORB itself (oracle/make_golden.py --orb, cv2 4.13.0), and the host-only geometry of the C ABI
(pyramid sizes, per-level quotas, INTER_LINEAR_EXACT taps) against the oracle.  No GPU."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import np_oracle as NO
from oracle import orb_oracle as OO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CFGS = {"default": {}, "n1000": dict(nfeatures=1000), "n200_l4_s15": dict(nfeatures=200, nlevels=4, scale_factor=1.5),
        "edge16_fast10": dict(edge_threshold=16, fast_threshold=10), "edge8": dict(edge_threshold=8)}


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def orb_golden():
    with open(os.path.join(GOLD, "orb_general.json")) as f:
        return json.load(f)


def golden_gray(case, small_clip, synth, cache={}):
    """Regenerate the gray frame of a golden case from its seed (checksum pinned in the fixture)."""
    key = (case["clip"], case["frame"])
    if key not in cache:
        name, i = key
        if name == "small":
            g = NO.bgr2gray(small_clip[i])
        elif name == "mid":
            g = NO.bgr2gray(synth.synth_clip(5, 270, 480, seed=3)[i])
        elif name == "hd":
            g = NO.bgr2gray(synth.synth_clip(3, 1080, 1920, seed=0)[i])
        elif name == "uhd":
            g = NO.bgr2gray(synth.synth_clip(1, 2160, 3840, seed=2)[i])
        elif name == "hd_resized_640x360":
            g = NO.bgr2gray(NO.resize_linear_u8(synth.synth_clip(3, 1080, 1920, seed=0)[i], 640, 360))
        else:
            g = np.random.default_rng(21).integers(0, 256, (200, 333), dtype=np.uint8)
        cache[key] = g
    assert _sha(cache[key]) == case["gray_sha"], "synthetic generator or gray/resize oracle drifted"
    return cache[key]


def digest(rows):
    """Same digest as oracle/make_golden.py orb_general: sorted (octave, response) list."""
    rr = np.array(sorted((r[0], float(np.float32(r[3]))) for r in rows), dtype=np.float64).reshape(-1, 2)
    key = np.concatenate([rr[:, 0].astype(np.int32).view(np.uint8), rr[:, 1].astype(np.float32).view(np.uint8)])
    return _sha(key)


def digest_keypoints(rows):
    """sha of the sorted (octave, pt.x, pt.y, size, angle, response) float32 table (make_golden.py)."""
    return _sha(np.array(sorted(rows), dtype=np.float32).reshape(-1, 6))


def test_orb_oracle_matches_cv2_fixtures(orb_golden, small_clip, synth):
    assert len(orb_golden["cases"]) >= 38
    for case in orb_golden["cases"]:
        gray = golden_gray(case, small_clip, synth)
        rows, per = OO.orb_detect(gray, **CFGS[case["cfg"]])
        assert per == case["per_level"], (case["clip"], case["frame"], case["cfg"])
        assert len(rows) == case["count"]
        assert digest(rows) == case["digest"], "Harris responses are not bit-identical to cv2's"
        # every cv2.KeyPoint field incl. the orientation
        assert digest_keypoints(OO.orb_keypoints(gray, **CFGS[case["cfg"]])) == case["digest_keypoints"], \
            (case["clip"], case["frame"], case["cfg"])


def test_orb_at_64x64_is_the_reference_path(small_clip):
    """At the reference's hard-wired size the general pipeline and the four-pixel shortcut agree."""
    from oracle import c_oracle as CO
    for f in small_clip:
        g = NO.bgr2gray(NO.resize_linear_u8(f, 64, 64))
        assert OO.orb_count(g) == CO.orb_count_64(g)


def test_orientation_known_answers():
    assert OO.umax_table() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert OO.fast_atan2(0, 0) == 0 and OO.fast_atan2(0, 5) == 0
    assert abs(float(OO.fast_atan2(3, 3)) - 45) < 0.01 and abs(float(OO.fast_atan2(-2, 0)) - 270) < 0.01
    ramp = np.tile(np.arange(64, dtype=np.uint8) * 3, (64, 1))          # brighter to the right: centroid at +x, angle 0
    assert OO.ic_angle(ramp, 32, 32) == 0
    assert abs(float(OO.ic_angle(np.ascontiguousarray(ramp.T), 32, 32)) - 90) < 0.01


def test_retain_best_keeps_ties():
    r = np.array([5, 3, 3, 3, 1, 9], np.float32)
    assert OO.retain_best(r, 2).tolist() == [True, False, False, False, False, True]
    assert OO.retain_best(r, 3).sum() == 5                     # the three 3s tie for third place
    assert OO.retain_best(r, 6).all() and OO.retain_best(r, 10).all()
    assert not OO.retain_best(r, 0).any()


@pytest.mark.parametrize("h,w", [(1080, 1920), (2160, 3840), (97, 131), (64, 64), (720, 1280), (333, 777)])
def test_abi_geometry_matches_oracle(vqa, h, w):
    from rtvqa_b200 import _native as N
    lw, lh, q = N.orb_describe(h, w)
    assert list(zip(lh, lw)) == OO.level_sizes(h, w)
    assert q == OO.level_quotas() == [109, 90, 75, 63, 52, 44, 36, 31]
    for l in range(1, 8):
        for sn, dn in ((lw[l - 1], lw[l]), (lh[l - 1], lh[l])):
            off, c1 = N.exact_taps(sn, dn)
            o, _, cc1 = OO.linear_exact_taps(sn, dn)
            assert np.array_equal(off, o) and np.array_equal(c1, cc1), (sn, dn)


@pytest.mark.parametrize("nf,nl,sf", [(1000, 8, 1.2), (200, 4, 1.5), (50, 12, 1.1), (500, 1, 1.2), (3, 8, 1.2), (0, 8, 1.2)])
def test_abi_quotas_for_other_configs(vqa, nf, nl, sf):
    from rtvqa_b200 import _native as N
    lw, lh, q = N.orb_describe(480, 640, N.orb_cfg(nf, sf, nl))
    assert q == OO.level_quotas(nf, nl, sf) and sum(q) >= nf - nl
    assert list(zip(lh, lw)) == OO.level_sizes(480, 640, nl, sf)


def test_abi_taps_up_and_down(vqa):
    from rtvqa_b200 import _native as N
    for sn, dn in [(10, 7), (7, 10), (1, 5), (5, 1), (100, 100), (1920, 1600), (3, 2)]:
        off, c1 = N.exact_taps(sn, dn)
        o, _, cc1 = OO.linear_exact_taps(sn, dn)
        assert np.array_equal(off, o) and np.array_equal(c1, cc1), (sn, dn)
    with pytest.raises(N.VqaError):
        N.exact_taps(0, 4)


def test_orb_size_knob_validation(vqa):
    from rtvqa_b200 import complexity_metrics as cm, video_processing as vp
    cm.set_orb_size((640, 360))
    assert cm.ORB_SIZE == (640, 360) and cm._orb_size(None) == (640, 360) and cm._orb_size((32, 16)) == (32, 16)
    cm.set_orb_size(None)
    assert cm.ORB_SIZE is None and cm._orb_size(None) is None
    with pytest.raises(ValueError):
        cm.set_orb_size((0, 5))
    base = dict(crf=23, resize_width=64, resize_height=64, frame_interval=10)
    vp.validate_config(dict(base))
    vp.validate_config(dict(base, orb_width=640, orb_height=360))
    for bad in (dict(orb_width=640), dict(orb_width=0, orb_height=5), dict(orb_width="a", orb_height=5)):
        with pytest.raises(ValueError):
            vp.validate_config(dict(base, **bad))
