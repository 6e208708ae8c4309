"""CPU-side checks: the C-ABI library loads and exports every symbol include/vqa_b200.h declares,
struct layouts agree between the header and the binding, the host mirrors keep the reference's
error behaviour / CSV contract, and the product path fails loudly without a GPU (no CPU fallback,
no route through oracle/)."""
import ctypes
import functools
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "real-time-video-quality-analysis_b200")
HDR = os.path.join(ROOT, "include", "vqa_b200.h")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol(vqa):
    from rtvqa_b200 import _native as N
    lib = N.load_library()
    declared = re.findall(r"VQA_API[^;(]*?\b(vqa_[a-z0-9_]+)\s*\(", open(HDR).read())
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in vqa_b200.h but not exported"
    assert sorted(set(declared)) == sorted(set(N.EXPORTS)), "binding and header disagree on the entry points"
    assert lib.vqa_abi_version() == N.ABI_VERSION == 3


def test_struct_layouts_match_header(vqa, tmp_path):
    from rtvqa_b200 import _native as N
    src = tmp_path / "layout.c"
    src.write_text('''
#include <stdio.h>
#include <stddef.h>
#include "vqa_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(vqa_frame_metrics), offsetof(vqa_frame_metrics, hist_entropy),
    offsetof(vqa_frame_metrics, color_entropy), offsetof(vqa_frame_metrics, dct_energy), offsetof(vqa_frame_metrics, motion),
    offsetof(vqa_frame_metrics, temporal_dct), offsetof(vqa_frame_metrics, orb_count), offsetof(vqa_frame_metrics, edge_count),
    offsetof(vqa_frame_metrics, gray_sq_sum));
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(vqa_fr_metrics), offsetof(vqa_fr_metrics, sse), offsetof(vqa_fr_metrics, mse),
    offsetof(vqa_fr_metrics, mse_avg), offsetof(vqa_fr_metrics, psnr), offsetof(vqa_fr_metrics, psnr_avg),
    offsetof(vqa_fr_metrics, ssim), offsetof(vqa_fr_metrics, ssim_all));
  printf("%zu %zu %zu\\n", sizeof(vqa_cfg), offsetof(vqa_cfg, orb_width), offsetof(vqa_cfg, orb_height));
  printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(vqa_orb_cfg), offsetof(vqa_orb_cfg, nfeatures), offsetof(vqa_orb_cfg, nlevels),
    offsetof(vqa_orb_cfg, edge_threshold), offsetof(vqa_orb_cfg, fast_threshold), offsetof(vqa_orb_cfg, scale_factor));
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(vqa_keypoint), offsetof(vqa_keypoint, x), offsetof(vqa_keypoint, y),
    offsetof(vqa_keypoint, size), offsetof(vqa_keypoint, angle), offsetof(vqa_keypoint, response), offsetof(vqa_keypoint, octave), offsetof(vqa_keypoint, lx), offsetof(vqa_keypoint, ly),
    offsetof(vqa_keypoint, fast_score));
  return 0; }''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    a, b, c, d, e = subprocess.check_output([str(exe)], text=True).strip().splitlines()
    f = N.FRAME_DTYPE
    assert [int(v) for v in a.split()] == [f.itemsize] + [f.fields[n][1] for n in f.names]
    g = N.FR_DTYPE
    assert [int(v) for v in b.split()] == [g.itemsize] + [g.fields[n][1] for n in g.names]
    assert [int(v) for v in c.split()] == [ctypes.sizeof(N.Cfg), N.Cfg.orb_width.offset, N.Cfg.orb_height.offset]
    assert [int(v) for v in d.split()] == [ctypes.sizeof(N.OrbCfg)] + [getattr(N.OrbCfg, n).offset for n, _ in N.OrbCfg._fields_]
    k = N.KEYPOINT_DTYPE
    assert [int(v) for v in e.split()] == [k.itemsize] + [k.fields[n][1] for n in k.names]


def test_product_package_never_imports_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), fn
                assert "vqa_oracle" not in text and "c_oracle" not in text, fn


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(vqa):
    from rtvqa_b200 import _native as N
    from rtvqa_b200 import complexity_metrics as cm
    with pytest.raises(N.VqaError, match="no CPU fallback"):
        N.Context(0)
    frame = np.zeros((32, 32, 3), np.uint8)
    with pytest.raises(N.VqaError):
        cm.process_dct_frame(frame, 16, 16)
    with pytest.raises(N.VqaError):
        cm.process_in_batches([frame], functools.partial(cm.process_edge_frame, resize_width=8, resize_height=8), 2)


def test_missing_library_is_an_error(vqa, monkeypatch):
    from rtvqa_b200 import _native as N
    monkeypatch.setattr(N, "_lib", None)
    monkeypatch.setattr(N, "SO_PATH", "/nonexistent/libvqa_b200.so")
    with pytest.raises(N.VqaError, match="no CPU fallback"):
        N.load_library()


def test_reference_error_conventions(vqa):
    from rtvqa_b200 import complexity_metrics as cm
    from rtvqa_b200 import video_processing as vp
    with pytest.raises(ValueError):
        cm.validate_video_path(123)
    with pytest.raises(ValueError):
        cm.validate_video_path("clip.mkv")
    assert cm.validate_video_path("a.mp4") == "video" and cm.validate_video_path("a.png") == "frame"
    assert cm.read_frame_pairs("/nonexistent/clip.mp4", 10) == []          # logs an error, returns []
    assert cm.extract_frame_timestamps("/nonexistent/clip.mp4", 10) == []
    assert cm.process_frame_complexity((None, np.zeros((4, 4, 3), np.uint8))) == 0.0
    with pytest.raises(TypeError, match="no CPU fallback"):
        cm.process_in_batches([1, 2], lambda f: 0, 2)
    for bad in ({"crf": 0, "resize_width": 64, "resize_height": 64}, {"crf": 23, "resize_width": 0, "resize_height": 64},
                {"crf": 23, "resize_width": 64, "resize_height": 64, "frame_interval": 0},
                {"crf": 23, "resize_width": 64, "resize_height": 64, "num_workers": "4"}):
        with pytest.raises(ValueError):
            vp.validate_config(bad)
    vp.validate_config({"crf": 23, "vmaf_model_path": None, "resize_width": 64, "resize_height": 64, "frame_interval": 10})
    with pytest.raises(FileNotFoundError):
        vp.process_video_and_extract_metrics("/nonexistent/in.mp4", {"crf": 23})
    with pytest.raises(FileNotFoundError):
        vp.load_config("/nonexistent/config.json")


def test_smooth_data_and_score(vqa, golden, monkeypatch):
    from rtvqa_b200 import complexity_metrics as cm
    np.testing.assert_allclose(cm.smooth_data(golden["ewm_in"], 0.8), golden["ewm_out"], rtol=1e-13)
    np.testing.assert_allclose(cm.smooth_data(golden["ewm_in"], 0.3), golden["ewm_out_a03"], rtol=1e-13)
    assert cm.smooth_data([]).shape == (0,)
    vals = (5.0, 2.55e7, 4.0, 0.5, 2500, 4.0, 5e6, 1.0)
    monkeypatch.setattr(cm, "calculate_average_scene_complexity", lambda *a, **k: vals)
    assert cm.calculate_scene_complexity_score("x.mp4", 64, 64) == pytest.approx(0.5)      # every metric at mid-range
    assert cm.normalize(3, 5, 5) == 0


def test_csv_contract(vqa, tmp_path):
    from rtvqa_b200 import video_processing as vp
    vals = tuple(float(i) for i in range(1, 9))      # motion, dct, hist, edge, orb, colour, tdct, fps
    cols = vp.complexity_columns(vals)
    # the reference's positional unpack (video_processing.py:235-242): values land under shifted names
    assert cols == {'Advanced Motion Complexity': 1.0, 'DCT Complexity': 2.0, 'Temporal DCT Complexity': 3.0,
                    'Histogram Complexity': 4.0, 'Edge Detection Complexity': 5.0, 'ORB Feature Complexity': 6.0,
                    'Color Histogram Complexity': 7.0, 'Framerate Variation': 8.0}
    fixed = vp.complexity_columns(vals, correct_column_order=True)
    assert fixed['Temporal DCT Complexity'] == 7.0 and fixed['Histogram Complexity'] == 3.0 and fixed['ORB Feature Complexity'] == 5.0
    csv_file = tmp_path / "video_quality_data.csv"
    row = {'Bitrate (kbps)': 4486, 'Resolution (px)': '1920x1080', 'Frame Rate (fps)': 30.0, 'CRF': 23, 'PSNR': 50.78,
           'SSIM': 0.994884, 'VMAF': 95.837165, **cols}
    vp.thread_safe_update_csv(row, str(csv_file))
    vp.thread_safe_update_csv(row, str(csv_file))
    lines = csv_file.read_text().splitlines()
    assert len(lines) == 3 and lines[0].split(",") == vp.CSV_COLUMNS          # README.md:71 header, written once
    assert lines[1].startswith("4486,1920x1080,30.0,23,50.78,0.994884,95.837165,1.0,2.0")
    # a missing metric is an empty field, as DataFrame.to_csv writes it (reference :59-65)
    vp.thread_safe_update_csv({**row, 'VMAF': float("nan"), 'PSNR': None}, str(csv_file))
    assert csv_file.read_text().splitlines()[3].startswith("4486,1920x1080,30.0,23,,0.994884,,1.0")


def test_smooth_data_follows_pandas_on_missing_observations(vqa):
    """pd.Series(x).ewm(alpha).mean() with NaN terms (reference :114-125): skipped, weights keep decaying, the value
    is carried forward, leading NaNs stay NaN; np.mean of it is what calculate_average_scene_complexity returns."""
    pd = pytest.importorskip("pandas")
    from rtvqa_b200 import complexity_metrics as cm
    rng = np.random.default_rng(5)
    for trial in range(12):
        x = rng.normal(size=33)
        x[rng.integers(0, 33, size=trial % 6)] = np.nan
        if trial % 4 == 0:
            x[0] = np.nan
        want = pd.Series(x).ewm(alpha=0.8).mean().to_numpy()
        np.testing.assert_allclose(cm.smooth_data(x, 0.8), want, rtol=1e-13, atol=0, equal_nan=True)
        if np.isnan(x).any():                                   # host path; NaN-free series go through the device reduction
            got = cm._smoothed_mean(x, 0.8)
            assert (np.isnan(got) and np.isnan(np.mean(want))) or got == pytest.approx(np.mean(want), rel=1e-13)


def test_stats_files_round_trip_through_the_reference_regexes(vqa, tmp_path):
    from rtvqa_b200 import _native as N
    from rtvqa_b200 import video_processing as vp
    rows = np.zeros(3, dtype=N.FR_DTYPE)
    rows["psnr_avg"] = [np.inf, 41.23456, 39.5]
    rows["psnr"] = [[np.inf] * 3, [40, 45, 46], [38, 44, 45]]
    rows["mse_avg"] = [0, 4.9, 7.3]
    rows["ssim_all"] = [1.0, 0.987654321, 0.95]
    rows["ssim"] = [[1, 1, 1], [0.98, 0.99, 0.995], [0.94, 0.96, 0.97]]
    p, s, v = tmp_path / "psnr.log", tmp_path / "ssim.log", tmp_path / "vmaf.json"
    vp.write_ffmpeg_stats(rows, str(p), str(s))
    v.write_text('{"pooled_metrics": {"vmaf": {"mean": 95.5}}}')
    m = vp.extract_metrics_from_logs(str(p), str(s), str(v), "in.mp4", 23, 4486, "1920x1080", 30.0)
    # an `inf` first frame is skipped by the reference's regex -> next frame's value
    assert m["PSNR"] == 41.23 and m["SSIM"] == 1.0 and m["VMAF"] == 95.5 and m["CRF"] == 23
    assert list(m)[:4] == ['Bitrate (kbps)', 'Resolution (px)', 'Frame Rate (fps)', 'CRF']


def test_drop_in_module_names_resolve_from_the_package_dir():
    """`import complexity_metrics` / `import video_processing` with the package dir on sys.path."""
    code = ("import sys; sys.path.insert(0, %r); import complexity_metrics as c, video_processing as v; "
            "print(c.calculate_average_scene_complexity.__name__, v.run_ffmpeg_metrics.__name__)" % PKG)
    out = subprocess.check_output([sys.executable, "-c", code], text=True, stderr=subprocess.DEVNULL)
    assert out.split() == ["calculate_average_scene_complexity", "run_ffmpeg_metrics"]


def test_product_build_reads_no_environment():
    """Every getenv of csrc/ sits inside `#ifdef VQA_AB` (the development build of the A/B notes): the product library's
    numerics and speed cannot be changed from the environment (round-1 finding: 14 knobs lived in the hot path)."""
    csrc = os.path.join(PKG, "csrc")
    bad = []
    for name in sorted(os.listdir(csrc)):
        if not name.endswith((".cu", ".cuh")):
            continue
        depth_ab, stack = 0, []
        for ln, line in enumerate(open(os.path.join(csrc, name)), 1):
            t = line.strip()
            if t.startswith(("#ifdef", "#ifndef", "#if ")):
                stack.append("VQA_AB" in t and t.startswith("#ifdef"))
            elif t.startswith("#else") and stack:
                stack[-1] = False                      # the #else branch of an #ifdef VQA_AB is product code
            elif t.startswith("#endif") and stack:
                stack.pop()
            if "getenv" in t and not t.startswith("//") and not any(stack):
                bad.append("%s:%d" % (name, ln))
    assert not bad, bad
