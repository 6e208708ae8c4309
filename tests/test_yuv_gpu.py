"""GPU parity of round 2's additions, through the C ABI: the yuv420p -> BGR conversion (bit-exact against the
cv2.VideoCapture fixtures and the oracle), the one-upload clip call vqa_analyze_clip_yuv420 (identical to the two
halves on the derived BGR frames), the vectorised PSNR/SSIM kernel on aligned and ragged planes, and the
library's own NCCL reduce (vqa_clip_reduce)."""
import hashlib
import json
import os
import socket

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import np_oracle as NO
from oracle import ref_port as RP

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def ctx(vqa):
    from rtvqa_b200 import _native as N
    c = N.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def yuv_golden():
    with open(os.path.join(GOLD, "yuv2bgr_cv2.json")) as f:
        return json.load(f)


def _planes(rng, n, h, w):
    return (rng.integers(0, 256, (n, h, w), dtype=np.uint8), rng.integers(0, 256, (n, h // 2, w // 2), dtype=np.uint8),
            rng.integers(0, 256, (n, h // 2, w // 2), dtype=np.uint8))


# ------------------------------------------------------------------ f4: yuv420p -> BGR (bit-exact)
def test_yuv2bgr_every_triple_matches_cv2(ctx, yuv_golden):
    from test_yuv_cpu import exhaustive_yuv_frame
    Y, U, V = exhaustive_yuv_frame()
    assert sha(ctx.debug_yuv2bgr(Y, U, V)) == yuv_golden["exhaustive_sha"]


def test_yuv2bgr_fixture_sizes_and_stored_case(ctx, yuv_golden):
    rng = np.random.default_rng(2024)                       # the generator's stream (oracle/make_golden.py --yuv2bgr)
    for case in yuv_golden["cases"]:
        Y, U, V = _planes(rng, case["n"], case["h"], case["w"])
        assert sha(Y) == case["y_sha"]
        got = np.stack([ctx.debug_yuv2bgr(Y[i], U[i], V[i]) for i in range(case["n"])])
        assert sha(got) == case["bgr_sha"], (case["h"], case["w"])
    z = np.load(os.path.join(GOLD, "yuv2bgr_cv2.npz"))
    for i in range(len(z["y"])):
        assert np.array_equal(ctx.debug_yuv2bgr(z["y"][i], z["u"][i], z["v"][i]), z["bgr"][i])


def test_yuv2bgr_odd_sizes_are_refused(ctx):
    from rtvqa_b200 import _native as N
    rng = np.random.default_rng(1)
    with pytest.raises((N.VqaError, TypeError)):
        ctx.analyze_clip_yuv420(_planes(rng, 1, 33, 48), None, 64, 64)


# ------------------------------------------------------------------ one upload, both halves
# (144, 192) and (72, 96) at native resolution take the fused ingest (planes -> gray + histograms, BGR only in registers, histogram
# moments for the DCT); (70, 102) at native resolution cannot (w % 8 != 0) and (.., 64, 64) must not: both go through BGR frames
@pytest.mark.parametrize("h,w,rw,rh", [(72, 96, 64, 64), (144, 192, 192, 144), (72, 96, 96, 72), (70, 102, 64, 64), (70, 102, 102, 70)])
def test_analyze_clip_yuv420_equals_the_two_halves(ctx, synth, h, w, rw, rh):
    import torch
    n = 7
    (ry, ru, rv), (dy, du, dv) = synth.synth_yuv_pairs(n, h, w, seed=4)
    bgr = NO.yuv420_to_bgr(dy, du, dv)                     # what cv2.VideoCapture decodes from the encode
    want_rows = ctx.complexity_frames(bgr, rw, rh)
    want_fr = ctx.psnr_ssim((dy, du, dv), (ry, ru, rv))
    rows, fr = ctx.analyze_clip_yuv420((dy, du, dv), (ry, ru, rv), rw, rh)
    for f in rows.dtype.names:
        if f != "motion":                                   # Farneback's frame sum uses double atomics: 1e-6, not bitwise
            assert np.array_equal(rows[f], want_rows[f], equal_nan=True), f
    np.testing.assert_allclose(rows["motion"][1:], want_rows["motion"][1:], rtol=1e-6)
    for f in fr.dtype.names:
        assert np.array_equal(fr[f], want_fr[f]), f
    # device-resident planes, and a range that starts inside the clip with the previous frame as halo
    dev = lambda planes: [torch.from_numpy(np.ascontiguousarray(p)).cuda() for p in planes]
    rows_d, fr_d = ctx.analyze_clip_yuv420(dev((dy, du, dv)), dev((ry, ru, rv)), rw, rh)
    for f in rows.dtype.names:
        if f not in ("motion",):                            # Farneback sums use float atomics: equal to 1e-6, not bitwise
            assert np.array_equal(rows_d[f], rows[f], equal_nan=True), f
    np.testing.assert_allclose(rows_d["motion"][1:], rows["motion"][1:], rtol=1e-6)
    assert np.array_equal(fr_d["sse"], fr["sse"]) and np.array_equal(fr_d["ssim_all"], fr["ssim_all"])
    a = 3
    halo = (dy[a - 1], du[a - 1], dv[a - 1])
    tail, fr_tail = ctx.analyze_clip_yuv420([p[a:] for p in (dy, du, dv)], [p[a:] for p in (ry, ru, rv)], rw, rh, halo_planes=halo)
    for f in ("edge_count", "orb_count", "hist_entropy", "color_entropy", "dct_energy", "temporal_dct", "gray_sq_sum"):
        assert np.array_equal(tail[f], rows[f][a:]), f
    np.testing.assert_allclose(tail["motion"], rows["motion"][a:], rtol=1e-6)
    assert np.array_equal(fr_tail["sse"], fr["sse"][a:])
    # complexity only (no reference planes)
    only, none = ctx.analyze_clip_yuv420((dy, du, dv), None, rw, rh)
    assert none is None and np.array_equal(only["edge_count"], rows["edge_count"])
    # and the rows are the oracle's rows of the decoded frames
    for i in range(n):
        assert int(rows["edge_count"][i]) == int(RP.o_edge(bgr[i], rw, rh))
        np.testing.assert_allclose(rows["hist_entropy"][i], RP.o_hist(bgr[i], rw, rh), rtol=2e-6)


def test_analyze_clip_yuv420_many_chunks_1080p(ctx, synth):
    """More frames than one device chunk at the bench size: chunk boundaries and the staging ring."""
    import torch
    from rtvqa_b200.synth_device import DeviceClipSynth
    syn = DeviceClipSynth(1080, 1920, 7, "cuda")
    ref, enc = syn.pairs(0, 60)
    rows_d, fr_d = ctx.analyze_clip_yuv420(enc, ref, 1920, 1080)
    rows_h, fr_h = ctx.analyze_clip_yuv420([p.cpu().numpy() for p in enc], [p.cpu().numpy() for p in ref], 1920, 1080)
    for f in ("edge_count", "orb_count", "gray_sq_sum", "hist_entropy", "color_entropy", "dct_energy"):
        assert np.array_equal(rows_d[f], rows_h[f]), f
    np.testing.assert_allclose(rows_d["motion"][1:], rows_h["motion"][1:], rtol=1e-6)
    np.testing.assert_allclose(rows_d["temporal_dct"][1:], rows_h["temporal_dct"][1:], rtol=1e-6)
    assert np.array_equal(fr_d["sse"], fr_h["sse"]) and np.array_equal(fr_d["ssim"], fr_h["ssim"])
    # three frames against the CPU oracle at full size
    e = [p[:3].cpu().numpy() for p in enc]
    r = [p[:3].cpu().numpy() for p in ref]
    bgr = NO.yuv420_to_bgr(*e)
    for i in range(3):
        assert int(rows_d["edge_count"][i]) == int(RP.o_edge(bgr[i], 1920, 1080))
        for c in range(3):
            assert int(fr_d["sse"][i, c]) == CO.plane_sse(e[c][i], r[c][i])
            assert fr_d["ssim"][i, c] == pytest.approx(CO.ssim_plane(e[c][i], r[c][i]), rel=1e-6)
    np.testing.assert_allclose(rows_d["motion"][1], RP.o_motion((bgr[1], bgr[0])), rtol=1e-4)


# ------------------------------------------------------------------ a13: both PSNR/SSIM kernels
@pytest.mark.parametrize("h,w", [(64, 64), (1080, 1920), (2160, 3840), (130, 96), (36, 1056), (66, 34), (8, 16)])
def test_psnr_ssim_vector_and_generic_kernels(ctx, h, w):
    """w % 32 == 0 takes k_psnr_ssim_v (chroma width % 16 == 0), other sizes the generic kernel; heights that are
    not multiples of 4 / 32 exercise the SSE tail rows and the partial last segment; 4K spans 8 strips."""
    import torch
    rng = np.random.default_rng(h * 7 + w)
    n = 2
    ref = _planes(rng, n, h, w)
    enc = tuple(np.clip(p.astype(np.int16) + rng.integers(-9, 10, p.shape), 0, 255).astype(np.uint8) for p in ref)
    got = ctx.psnr_ssim(enc, ref)
    dev = ctx.psnr_ssim([torch.from_numpy(p).cuda() for p in enc], [torch.from_numpy(p).cuda() for p in ref])
    for f in got.dtype.names:
        assert np.array_equal(got[f], dev[f]), f            # deterministic: host-staged == device-resident, bit for bit
    for i in range(n):
        for c in range(3):
            assert int(got["sse"][i, c]) == CO.plane_sse(enc[c][i], ref[c][i])
            bw, bh = ref[c].shape[2] >> 2, ref[c].shape[1] >> 2
            if bw > 1 and bh > 1:
                assert got["ssim"][i, c] == pytest.approx(CO.ssim_plane(enc[c][i], ref[c][i]), rel=1e-6)
    again = ctx.psnr_ssim(enc, ref)
    assert again.tobytes() == got.tobytes()                 # run-to-run identical (no float atomics on this path)


def test_psnr_ssim_many_frames_chunked_upload(ctx):
    rng = np.random.default_rng(9)
    n, h, w = 150, 64, 96                                   # > one 64-pair staging chunk
    ref = _planes(rng, n, h, w)
    enc = tuple(np.clip(p.astype(np.int16) + rng.integers(-3, 4, p.shape), 0, 255).astype(np.uint8) for p in ref)
    got = ctx.psnr_ssim(enc, ref)
    for i in (0, 63, 64, 127, 128, 149):
        for c in range(3):
            assert int(got["sse"][i, c]) == CO.plane_sse(enc[c][i], ref[c][i])
            assert got["ssim"][i, c] == pytest.approx(CO.ssim_plane(enc[c][i], ref[c][i]), rel=1e-6)


# ------------------------------------------------------------------ Canny: repeatability under contention
def test_canny_repeatable_under_contention(ctx):
    """compute-sanitizer is closed on the GPU pool (profiles/r02_sanitizer_closed.txt), so the union-find of the
    hysteresis (shared-memory atomicMin unions against volatile finds) is exercised the other way round: dense
    inputs with thousands of merges per tile, 40 repetitions, every one bit-identical to the CPU oracle."""
    from rtvqa_b200 import _native as N
    rng = np.random.default_rng(5)
    h, w = 272, 480
    g = (rng.random((h, w)) < 0.35).astype(np.uint8) * rng.integers(90, 256, (h, w)).astype(np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    g2 = (127 + 120 * np.sin(np.hypot(xx - 240, yy - 136) / 1.7)).astype(np.uint8)
    for img in (g, g2):
        want_n, want_map = CO.canny_count(img, want_map=True)
        bgr = np.repeat(img[..., None], 3, axis=2)[None].repeat(8, axis=0)
        for _ in range(5):
            assert np.array_equal(ctx.debug_canny(img), want_map)
            rows = ctx.complexity_frames(bgr, w, h, N.M_EDGE)
            assert rows["edge_count"].tolist() == [int(want_n)] * 8


# ------------------------------------------------------------------ e: the library's own NCCL reduce
def test_clip_reduce_single_rank_and_errors(ctx):
    from rtvqa_b200 import _native as N
    with pytest.raises(N.VqaError, match="no communicator"):
        ctx.clip_reduce(np.ones(3), np.ones(2, np.int64))
    ctx.comm_init(0, 1, N.comm_unique_id())
    p, i = ctx.clip_reduce(np.array([[1.5, -2.25], [1e300, 3.0]]), np.array([7, -3, 2 ** 52], np.int64))
    assert p.tolist() == [[1.5, -2.25], [1e300, 3.0]] and i.tolist() == [7, -3, 2 ** 52] and i.dtype == np.int64
    with pytest.raises(N.VqaError, match="too large"):
        ctx.clip_reduce(np.zeros(1), np.array([2 ** 53], np.int64))
    ctx.comm_destroy()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist
    import rtvqa_b200  # noqa: F401
    from rtvqa_b200 import _native as N
    from rtvqa_b200 import sharding as SH
    from rtvqa_b200.synth_device import DeviceClipSynth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ctx = N.get_context(rank)
    ctx.use_torch_stream()
    SH.init_context_comm(ctx)
    k, h, w = 12, 72, 96
    a, b = SH.shard_range(k, rank, world)
    syn = DeviceClipSynth(h, w, 3, f"cuda:{rank}")
    ref, enc = syn.pairs(a, b - a)
    halo = [p.contiguous() for p in syn.pair(a - 1)[1]] if a > 0 else None
    # the halo also travels over the library's grouped send/recv: it must equal the locally generated frame
    pack = torch.empty(h * w * 3 // 2, dtype=torch.uint8, device=f"cuda:{rank}")
    last = torch.cat([enc[c][-1].reshape(-1) for c in range(3)]).contiguous()
    ctx.halo_exchange(last if rank + 1 < world else None, pack if rank > 0 else None)
    torch.cuda.synchronize()
    halo_ok = True if rank == 0 else bool(torch.equal(pack, torch.cat([p.reshape(-1) for p in halo])))
    rows, fr = ctx.analyze_clip_yuv420(enc, ref, 64, 64, halo_planes=halo)
    partials = SH.local_partials(rows, a, k, 0.8, ctx.ewm_partial)
    lo = max(1 - a, 0)
    ints = np.array([int(rows["edge_count"][lo:].sum()), int(rows["orb_count"][lo:].sum()), len(rows)], dtype=np.int64)
    partials, ints = SH.reduce_partials(partials, ints, ctx=ctx)              # vqa_clip_reduce
    oks = [None] * world
    dist.all_gather_object(oks, halo_ok)
    if rank == 0:
        np.savez(out_path, p=partials, i=ints, halo_ok=all(oks))
    dist.barrier()
    ctx.comm_destroy()
    dist.destroy_process_group()


def test_two_rank_nccl_clip_reduce(vqa, ctx, tmp_path):
    """Frame-range sharding over 2 GPUs closed by vqa_clip_reduce (ONE ncclAllReduce inside the library) == the
    single-GPU pass over the whole clip; needs a 2-GPU box (skipped on the 1-GPU test tier)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from rtvqa_b200 import sharding as SH
    from rtvqa_b200.synth_device import DeviceClipSynth
    out = str(tmp_path / "nccl.npz")
    mp.spawn(_nccl_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    z = np.load(out)
    k = 12
    ref, enc = DeviceClipSynth(72, 96, 3, "cuda:0").pairs(0, k)
    rows, _ = ctx.analyze_clip_yuv420(enc, ref, 64, 64)
    want = SH.local_partials(rows, 0, k, 0.8, ctx.ewm_partial)
    np.testing.assert_allclose(z["p"], want, rtol=1e-12)
    assert z["i"].tolist() == [int(rows["edge_count"][1:].sum()), int(rows["orb_count"][1:].sum()), k] and bool(z["halo_ok"])
