"""The oracle (oracle/np_oracle.py + oracle/c/vqa_oracle.c) against fixtures produced by the
UNMODIFIED reference (oracle/make_golden.py, cv2 4.13.0).  CPU only."""
import hashlib

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import np_oracle as NO
from oracle import ref_port as RP


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _per_frame(clip, rw, rh):
    grays = [NO.resize_linear_u8(NO.bgr2gray(f), rw, rh) for f in clip]
    return dict(
        dct=[RP.o_dct(f, rw, rh) for f in clip],
        hist=[RP.o_hist(f, rw, rh) for f in clip],
        color=[RP.o_color(f, rw, rh) for f in clip],
        edge=[int(RP.o_edge(f, rw, rh)) for f in clip],
        orb=[RP.o_orb(f) for f in clip],
        motion=[float(RP.o_motion((clip[i], clip[i - 1]))) for i in range(1, len(clip))],
        tdct=[RP.o_tdct(grays[i - 1], grays[i], rw, rh) for i in range(1, len(clip))],
    )


def _check(got, want):
    assert got["edge"] == want["edge"]            # integer: bit-exact
    assert got["orb"] == want["orb"]
    np.testing.assert_allclose(got["hist"], want["hist"], rtol=2e-6)
    np.testing.assert_allclose(got["color"], want["color"], rtol=2e-6)
    np.testing.assert_allclose(got["dct"], want["dct"], rtol=1e-5)      # cv2.dct is float32
    np.testing.assert_allclose(got["tdct"], want["tdct"], rtol=1e-5)
    np.testing.assert_allclose(got["motion"], want["motion"], rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("key,rw,rh", [("small_64", 64, 64), ("small_native", 128, 96),
                                        ("small_odd", 100, 37), ("small_up", 160, 120)])
def test_small_clip_per_frame(golden, small_clip, key, rw, rh):
    assert _sha(small_clip) == golden["small_sha"]
    _check(_per_frame(small_clip, rw, rh), golden[key])


def test_mid_clip_per_frame(golden, synth):
    mid = synth.synth_clip(5, 270, 480, seed=3)
    assert _sha(mid) == golden["mid_sha"], "synthetic generator drifted"
    _check(_per_frame(mid, 480, 270), golden["mid_native"])
    _check(_per_frame(mid, 64, 64), golden["mid_64"])


@pytest.mark.parametrize("key,rw,rh,interval", [("small_avg_i1_64", 64, 64, 1), ("small_avg_i3_64", 64, 64, 3),
                                                 ("small_avg_i1_native", 128, 96, 1)])
def test_average_scene_complexity(golden, small_clip, key, rw, rh, interval):
    got = RP.average_scene_complexity(small_clip, rw, rh, frame_interval=interval)
    want = golden[key]
    np.testing.assert_allclose(got, want, rtol=1e-4)
    assert got[3] == pytest.approx(want[3], rel=1e-12)      # edge counts -> exact series
    assert got[4] == pytest.approx(want[4], rel=1e-12)
    assert got[7] == pytest.approx(want[7], rel=1e-12)


def test_ewm_and_fps(golden):
    np.testing.assert_allclose(NO.ewm_mean(golden["ewm_in"], 0.8), golden["ewm_out"], rtol=1e-13)
    np.testing.assert_allclose(NO.ewm_mean(golden["ewm_in"], 0.3), golden["ewm_out_a03"], rtol=1e-13)
    got = [NO.process_frame_interval_for_parallel(tuple(p)) for p in golden["fps_pairs"]]
    assert got == golden["fps_out"]


def test_sampling_semantics():
    # N=300, I=10 -> 30 samples (9,19,...), 29 pairs, 30 timestamps
    idx = NO.sampled_indices(300, 10)
    assert idx[0] == 9 and idx[-1] == 299 and len(idx) == 30
    assert len(NO.timestamp_indices(300, 10)) == 30 and NO.timestamp_indices(300, 10)[1] == 10
    assert len(NO.timestamp_indices(301, 10)) == 31


def test_known_answers_zero_frame():
    z = np.zeros((48, 64, 3), np.uint8)
    assert RP.o_dct(z, 64, 64) == 0.0
    assert float(RP.o_hist(z, 64, 64)) == 0.0
    assert float(RP.o_color(z, 64, 64)) == pytest.approx(0.0, abs=1e-6)
    assert RP.o_edge(z, 64, 64) == 0 and RP.o_orb(z) == 0
    assert float(RP.o_motion((z, z))) == 0.0
    assert RP.o_motion((None, z)) == 0.0


def test_psnr_ssim_known_answers(synth):
    (ry, ru, rv), (dy, du, dv) = synth.synth_yuv_pairs(2, 72, 96, seed=1)
    same = RP.psnr_ssim_frames((ry, ru, rv), (ry, ru, rv))
    assert np.all(np.isinf(same["psnr_avg"])) and np.allclose(same["ssim_all"], 1.0)
    r = RP.psnr_ssim_frames((dy, du, dv), (ry, ru, rv))
    assert np.all((r["psnr_avg"] > 30) & (r["psnr_avg"] < 60)) and np.all((r["ssim_all"] > 0.8) & (r["ssim_all"] < 1))
    # C and NumPy restatements of vf_psnr / vf_ssim agree
    for i in range(2):
        a = NO.psnr_frame((dy[i], du[i], dv[i]), (ry[i], ru[i], rv[i]))
        b = NO.ssim_frame((dy[i], du[i], dv[i]), (ry[i], ru[i], rv[i]))
        assert a["psnr_avg"] == pytest.approx(r["psnr_avg"][i], rel=1e-12)
        assert b["ssim_all"] == pytest.approx(r["ssim_all"][i], rel=1e-6)
    # ssim constants (vf_ssim.c): c1 = .01^2*255^2*64, c2 = .03^2*255^2*64*63
    assert int(.01 * .01 * 255 * 255 * 64 + .5) == 416 and int(.03 * .03 * 255 * 255 * 64 * 63 + .5) == 235963


def test_config1_reference_case(golden, synth):
    """BASELINE.json configs[0] -- the reference's own CPU-runnable case (300 x 1080p, frame_interval 10,
    resize 64x64): the oracle port against the 8-tuple the UNMODIFIED reference returned for it."""
    import numpy as np
    from helpers import SparseClip, config1_sampled_frames
    frames = config1_sampled_frames(synth)
    idx = sorted(frames)
    assert idx == list(range(9, 300, 10))
    assert _sha(np.stack([frames[i] for i in idx])) == golden["c1_sha_sampled"], "synthetic generator drifted"
    got = RP.average_scene_complexity(SparseClip(300, frames), 64, 64, frame_interval=10, workers=1)   # no fork inside a threaded pytest process
    want = golden["c1_avg"]
    np.testing.assert_allclose(got, want, rtol=1e-4)
    assert got[3] == pytest.approx(want[3], rel=1e-12) and got[4] == pytest.approx(want[4], rel=1e-12)   # edge, ORB: exact series
    assert got[7] == pytest.approx(3.0000000000000004, rel=1e-15)          # README.md:72 framerate of a CFR 30 fps clip at I = 10


def test_psnr_planes_against_cv2_fixture(synth):
    """a13, PSNR half: per-plane PSNR of the oracle against cv2.PSNR values (tests/golden/psnr_cv2.json).
    SSIM and the plane weighting remain on known answers only (no FFmpeg in the image: parity unpinned)."""
    import json, os
    from conftest import GOLD
    with open(os.path.join(GOLD, "psnr_cv2.json")) as f:
        g = json.load(f)
    cache = {}
    for case in g["cases"]:
        key = (case["n"], case["h"], case["w"], case["seed"])
        if key not in cache:
            cache[key] = synth.synth_yuv_pairs(*key[:3], seed=key[3])
        ref, dist = cache[key]
        i = case["frame"]
        assert [_sha(dist[c][i]) for c in range(3)] == case["plane_sha"]
        got = RP.psnr_ssim_frames(tuple(p[i:i + 1] for p in dist), tuple(p[i:i + 1] for p in ref))
        psnr = 10.0 * np.log10(255.0 * 255.0 / got["mse"][0])
        np.testing.assert_allclose(psnr, case["psnr"], rtol=1e-12)
