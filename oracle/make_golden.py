"""Generate tests/golden/* by running the UNMODIFIED reference (imported read-only from
/root/reference, cv2 from the image) on deterministic synthetic frames.

Run in the build container only:   python oracle/make_golden.py
The GPU box has no /root/reference; tests there read the committed fixtures.

Only two reference functions are monkey-patched, and only to serve in-memory frames with the
reference's own sampling: ``read_frame_pairs`` and
``extract_frame_timestamps`` (complexity_metrics.py:38-111).
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def _load_synth():
    p = os.path.join(ROOT, "real-time-video-quality-analysis_b200", "synth.py")
    spec = importlib.util.spec_from_file_location("vqa_synth", p)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _load_reference():
    sys.path.insert(0, "/root/reference")
    import complexity_metrics as ref  # noqa: E402  (the real thing)
    return ref


def patch_readers(ref, clip, fps=30.0):
    """Serve `clip` through the reference's two VideoCapture readers, sampling as they do."""
    n = len(clip)

    def read_frame_pairs(video_path, frame_interval=10):
        ref.validate_video_path(video_path)
        pairs, prev = [], None
        for count in range(1, n + 1):              # counter incremented before the test (:102-103)
            if count % frame_interval == 0:
                f = clip[count - 1]
                if prev is not None:
                    pairs.append((f, prev))
                prev = f
        return pairs

    def extract_frame_timestamps(video_path, frame_interval=10):
        ref.validate_video_path(video_path)
        return [1000.0 * i / fps for i in range(n) if i % frame_interval == 0]   # test before increment (:65-69)

    ref.read_frame_pairs = read_frame_pairs
    ref.extract_frame_timestamps = extract_frame_timestamps


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def per_frame(ref, clip, rw, rh):
    import cv2
    out = dict(
        dct=[float(ref.process_dct_frame(f, rw, rh)) for f in clip],
        hist=[float(ref.process_histogram_frame(f, rw, rh)) for f in clip],
        color=[float(ref.process_color_histogram_frame(f, rw, rh)) for f in clip],
        edge=[int(ref.process_edge_frame(f, rw, rh)) for f in clip],
        orb=[int(ref.process_orb_frame_for_parallel(f)) for f in clip],
        motion=[float(ref.process_frame_complexity((clip[i], clip[i - 1]))) for i in range(1, len(clip))],
    )
    grays = [cv2.resize(cv2.cvtColor(f, cv2.COLOR_BGR2GRAY), (rw, rh)) for f in clip]
    out["tdct"] = [float(ref.process_temporal_dct_frame(grays[i - 1], grays[i], rw, rh)) for i in range(1, len(clip))]
    return out


def main():
    import cv2
    import pandas as pd
    os.makedirs(GOLD, exist_ok=True)
    S = _load_synth()
    ref = _load_reference()
    meta = dict(cv2=cv2.__version__, numpy=np.__version__, pandas=pd.__version__,
                reference="/root/reference/complexity_metrics.py (unmodified, CPU branch)")

    # ---- small clip, stored explicitly --------------------------------------------------
    small = S.synth_clip(12, 96, 128, seed=7)
    g = dict(meta=meta, small_sha=sha(small))
    g["small_64"] = per_frame(ref, small, 64, 64)
    g["small_native"] = per_frame(ref, small, 128, 96)
    g["small_odd"] = per_frame(ref, small, 100, 37)
    g["small_up"] = per_frame(ref, small, 160, 120)
    patch_readers(ref, small)
    for interval in (1, 3):
        g[f"small_avg_i{interval}_64"] = [float(v) for v in ref.calculate_average_scene_complexity(
            "synthetic.mp4", 64, 64, frame_interval=interval, num_workers=2)]
    g["small_avg_i1_native"] = [float(v) for v in ref.calculate_average_scene_complexity(
        "synthetic.mp4", 128, 96, frame_interval=1, num_workers=2)]
    np.savez_compressed(os.path.join(GOLD, "small_clip.npz"), clip=small)

    # ---- mid-size frames (regenerated from the seed in tests; checksum pinned) -----------
    mid = S.synth_clip(5, 270, 480, seed=3)
    g["mid_sha"] = sha(mid)
    g["mid_native"] = per_frame(ref, mid, 480, 270)
    g["mid_64"] = per_frame(ref, mid, 64, 64)

    # ---- 1080p: 3 frames, full-res metrics (BASELINE.json config 2 shape) ----------------
    hd = S.synth_clip(3, 1080, 1920, seed=0)
    g["hd_sha"] = sha(hd)
    g["hd_native"] = per_frame(ref, hd, 1920, 1080)

    # ---- config 1 (reference CPU case): 300 x 1080p, I=10, 64x64 -------------------------
    c1 = S.synth_clip(300, 1080, 1920, seed=0)
    g["c1_sha_sampled"] = sha(c1[9::10])
    patch_readers(ref, c1)
    g["c1_avg"] = [float(v) for v in ref.calculate_average_scene_complexity(
        "synthetic.mp4", 64, 64, frame_interval=10, num_workers=os.cpu_count())]

    # ---- smoothing / framerate known answers ---------------------------------------------
    rng = np.random.default_rng(11)
    xs = rng.normal(size=17).tolist()
    g["ewm_in"] = xs
    g["ewm_out"] = [float(v) for v in ref.smooth_data(xs, 0.8)]
    g["ewm_out_a03"] = [float(v) for v in ref.smooth_data(xs, 0.3)]
    g["fps_pairs"] = [[0.0, 333.3333333333333], [100.0, 100.0], [50.0, 10.0], [0.0, 1000.0 / 30.0]]
    g["fps_out"] = [float(ref.process_frame_interval_for_parallel(tuple(p))) for p in g["fps_pairs"]]
    with open(os.path.join(GOLD, "reference_outputs.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", GOLD)


def orb_general(gray, **kw):
    """Keypoints of cv2.ORB_create(**kw).detectAndCompute(gray, None) -- the call the reference makes at
    complexity_metrics.py:385-387, at sizes other than its hard-wired 64x64 -- as per-level counts and
    a digest of the sorted (octave, response) list."""
    import cv2
    orb = cv2.ORB_create(**kw)
    kps, _ = orb.detectAndCompute(gray, None)
    nl = kw.get("nlevels", 8)
    per = [sum(1 for k in kps if k.octave == l) for l in range(nl)]
    rr = np.array(sorted((k.octave, float(np.float32(k.response))) for k in kps), dtype=np.float64).reshape(-1, 2)
    key = np.concatenate([rr[:, 0].astype(np.int32).view(np.uint8), rr[:, 1].astype(np.float32).view(np.uint8)])
    # every cv2.KeyPoint field: (octave, pt.x, pt.y, size, angle, response), sorted, float32 bit patterns
    full = np.array(sorted((k.octave, k.pt[0], k.pt[1], k.size, k.angle, k.response) for k in kps), dtype=np.float32).reshape(-1, 6)
    return dict(count=len(kps), per_level=per, digest=sha(key), digest_keypoints=sha(full))


ORB_CONFIGS = {
    "default": {},
    "n1000": dict(nfeatures=1000),
    "n200_l4_s15": dict(nfeatures=200, nlevels=4, scaleFactor=1.5),
    "edge16_fast10": dict(edgeThreshold=16, fastThreshold=10),
    "edge8": dict(edgeThreshold=8),                      # orientation patches reach into the reflect-101 border
}


def main_orb():
    """tests/golden/orb_general.json: cv2's ORB on the synthetic frames."""
    import cv2
    S = _load_synth()
    g = dict(meta=dict(cv2=cv2.__version__, source="cv2.ORB_create(**cfg).detectAndCompute(gray, None)"), cases=[])
    small = np.load(os.path.join(GOLD, "small_clip.npz"))["clip"]
    mid = S.synth_clip(5, 270, 480, seed=3)
    hd = S.synth_clip(3, 1080, 1920, seed=0)
    rng = np.random.default_rng(21)
    noise = rng.integers(0, 256, (200, 333), dtype=np.uint8)
    frames = [("small", i, cv2.cvtColor(small[i], cv2.COLOR_BGR2GRAY)) for i in (0, 5, 11)]
    frames += [("mid", i, cv2.cvtColor(mid[i], cv2.COLOR_BGR2GRAY)) for i in (0, 4)]
    frames += [("hd", 1, cv2.cvtColor(hd[1], cv2.COLOR_BGR2GRAY))]
    frames += [("noise_200x333_seed21", 0, noise)]
    uhd = S.synth_clip(1, 2160, 3840, seed=2)                    # BASELINE.json config 4 frame size
    frames += [("uhd", 0, cv2.cvtColor(uhd[0], cv2.COLOR_BGR2GRAY))]
    # resize -> gray order of the orb_size knob: gray(resize(frame, (w, h))), INTER_LINEAR
    frames += [("hd_resized_640x360", 1, cv2.cvtColor(cv2.resize(hd[1], (640, 360)), cv2.COLOR_BGR2GRAY))]
    for name, idx, gray in frames:
        for cname, cfg in ORB_CONFIGS.items():
            if (name == "hd" and cname not in ("default", "n1000")) or (name == "uhd" and cname != "default"):
                continue
            g["cases"].append(dict(clip=name, frame=idx, shape=list(gray.shape), gray_sha=sha(gray), cfg=cname,
                                   **orb_general(gray, **cfg)))
    with open(os.path.join(GOLD, "orb_general.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote orb_general.json:", len(g["cases"]), "cases")


def main_fr():
    """tests/golden/psnr_cv2.json: per-plane PSNR of synthetic yuv420p pairs from an independent library
    (cv2.PSNR).  FFmpeg itself is not in the image, so this pins the PSNR half of a13 only as far as
    10*log10(255^2 / mse) per plane goes; the area-weighted average and SSIM stay on known answers."""
    import cv2
    S = _load_synth()
    g = dict(meta=dict(cv2=cv2.__version__, source="cv2.PSNR(main_plane, ref_plane) (R = 255)"), cases=[])
    for (n, h, w, seed) in ((3, 72, 96, 2), (2, 270, 480, 1), (1, 1080, 1920, 1)):
        ref, dist = S.synth_yuv_pairs(n, h, w, seed=seed)
        for i in range(n):
            g["cases"].append(dict(n=n, h=h, w=w, seed=seed, frame=i,
                                   plane_sha=[sha(dist[c][i]) for c in range(3)],
                                   psnr=[float(cv2.PSNR(dist[c][i], ref[c][i])) for c in range(3)]))
    with open(os.path.join(GOLD, "psnr_cv2.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote psnr_cv2.json:", len(g["cases"]), "cases")


def write_y4m(path, Y, U, V):
    """Minimal yuv4mpeg2 writer (yuv420p): the one container cv2's bundled libavformat demuxes that carries
    raw planes, so cv2.VideoCapture.read() returns exactly swscale(yuv420p -> bgr24) of the planes given."""
    with open(path, "wb") as f:
        f.write(b"YUV4MPEG2 W%d H%d F30:1 Ip A1:1 C420jpeg\n" % (Y.shape[2], Y.shape[1]))
        for i in range(len(Y)):
            f.write(b"FRAME\n")
            f.write(Y[i].tobytes())
            f.write(U[i].tobytes())
            f.write(V[i].tobytes())


def read_video(path):
    import cv2
    cap = cv2.VideoCapture(path)
    out = []
    while True:
        ok, a = cap.read()
        if not ok:
            break
        out.append(a)
    cap.release()
    return np.stack(out)


def exhaustive_yuv_frame():
    """One 4096x4096 yuv420p frame in which every (Y,U,V) triple occurs exactly once."""
    cidx = np.arange(2048 * 2048).reshape(2048, 2048)
    U = ((cidx >> 6) & 255).astype(np.uint8)
    V = ((cidx >> 14) & 255).astype(np.uint8)
    s = cidx & 63
    Y = np.zeros((4096, 4096), np.uint8)
    Y[0::2, 0::2] = (4 * s).astype(np.uint8)
    Y[0::2, 1::2] = (4 * s + 1).astype(np.uint8)
    Y[1::2, 0::2] = (4 * s + 2).astype(np.uint8)
    Y[1::2, 1::2] = (4 * s + 3).astype(np.uint8)
    return Y, U, V


def main_yuv2bgr():
    """tests/golden/yuv2bgr_cv2.{json,npz}: BGR frames cv2.VideoCapture decodes from yuv4mpeg files whose planes
    we chose: (1) sha256 of the decode of the exhaustive frame (all 2^24 triples), (2) sha256 of
    decodes of random planes at even sizes from 2x2 to 1080p, (3) one small random case stored in full."""
    import tempfile
    import cv2
    g = dict(meta=dict(cv2=cv2.__version__, source="cv2.VideoCapture(<yuv4mpeg2 C420jpeg file>).read()",
                       note="odd frame sizes take another swscale path and are not covered"), cases=[])
    tmp = tempfile.mkdtemp()
    path = os.path.join(tmp, "t.y4m")
    Y, U, V = exhaustive_yuv_frame()
    write_y4m(path, Y[None], U[None], V[None])
    g["exhaustive_sha"] = sha(read_video(path)[0])
    rng = np.random.default_rng(2024)
    for (h, w) in ((2, 2), (4, 6), (16, 8), (18, 10), (50, 34), (98, 102), (144, 192), (146, 198), (270, 482), (1080, 1920)):
        n = 2 if h < 1000 else 1
        Yr = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
        Ur = rng.integers(0, 256, (n, h // 2, w // 2), dtype=np.uint8)
        Vr = rng.integers(0, 256, (n, h // 2, w // 2), dtype=np.uint8)
        write_y4m(path, Yr, Ur, Vr)
        got = read_video(path)
        assert got.shape == (n, h, w, 3)
        g["cases"].append(dict(h=h, w=w, n=n, seed=2024, y_sha=sha(Yr), u_sha=sha(Ur), v_sha=sha(Vr), bgr_sha=sha(got)))
        if (h, w) == (146, 198):
            np.savez_compressed(os.path.join(GOLD, "yuv2bgr_cv2.npz"), y=Yr, u=Ur, v=Vr, bgr=got)
    with open(os.path.join(GOLD, "yuv2bgr_cv2.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote yuv2bgr_cv2.json:", len(g["cases"]), "cases, exhaustive sha", g["exhaustive_sha"][:16])


if __name__ == "__main__":
    if "--fr" in sys.argv:
        main_fr()
    elif "--yuv2bgr" in sys.argv:
        main_yuv2bgr()
    elif "--orb" in sys.argv:
        main_orb()
    else:
        main()
