"""process_video_and_extract_metrics end to end (reference video_processing.py:180-267): encode ->
PSNR / SSIM / VMAF stats files -> ffprobe -> scene complexity of the ENCODED file -> one CSV row with the README's
15 columns.  The image has no ffmpeg / ffprobe, so test doubles of the two executables stand on PATH
(tests/helpers.py); the decoded frames they hand out are real (OpenCV's mp4v codec).

CPU variant: the two device entry points are replaced by the oracle, which checks the orchestration, the stats
files, the reference's parsing quirks and the CSV contract.  GPU variant: nothing is replaced."""
import csv
import os

import numpy as np
import pytest

from oracle import ref_port as RP

cv2 = pytest.importorskip("cv2")
CFG = {"crf": 23, "vmaf_model_path": None, "resize_width": 64, "resize_height": 64, "frame_interval": 3}


@pytest.fixture()
def cli(tmp_path, monkeypatch, small_clip):
    from helpers import install_fake_ffmpeg
    bin_dir = tmp_path / "bin"
    bin_dir.mkdir()
    tool = install_fake_ffmpeg(bin_dir)
    monkeypatch.setenv("PATH", str(bin_dir) + os.pathsep + os.environ["PATH"])
    monkeypatch.chdir(tmp_path)                                   # the CSV lands in the working directory (:262)
    src = str(tmp_path / "input.mp4")
    h, w = small_clip.shape[1:3]
    wr = cv2.VideoWriter(src, cv2.VideoWriter_fourcc(*"mp4v"), 30.0, (w, h))
    if not wr.isOpened():
        pytest.skip("no mp4v writer in this OpenCV build")
    for f in small_clip:
        wr.write(f)
    wr.release()
    return dict(src=src, tool=tool, tmp=tmp_path, stats_before=_stats_files())


def _stats_files():
    import re
    import tempfile
    return sorted(p for p in os.listdir(tempfile.gettempdir()) if re.match(r"(psnr|ssim|vmaf)_[0-9a-f]{32}\.(log|json)$", p))


def _expected(cli):
    """What the reference pipeline yields for these files, from the oracle."""
    from helpers import yuv420_planes
    tool, tmp = cli["tool"], cli["tmp"]
    enc = str(tmp / "expect_encoded.mp4")
    tool["encode"](cli["src"], enc)
    ref_frames, enc_frames = tool["decode"](cli["src"]), tool["decode"](enc)
    fr = RP.psnr_ssim_frames(yuv420_planes(enc_frames), yuv420_planes(ref_frames))
    vals = RP.average_scene_complexity(np.stack(enc_frames), CFG["resize_width"], CFG["resize_height"],
                                       frame_interval=CFG["frame_interval"])
    return fr, vals, enc_frames


def _check_csv(cli, fr, vals, rtol):
    from rtvqa_b200 import video_processing as vp
    with open(cli["tmp"] / "video_quality_data.csv") as f:
        rows = list(csv.reader(f))
    assert rows[0] == vp.CSV_COLUMNS and len(rows) == 2                      # README.md:71, header once
    row = dict(zip(rows[0], rows[1]))
    assert row["Bitrate (kbps)"] == "1234" and row["Resolution (px)"] == "128x96" and float(row["Frame Rate (fps)"]) == 30.0
    assert row["CRF"] == "23" and float(row["VMAF"]) == 93.25
    first = next(v for v in fr["psnr_avg"] if np.isfinite(v))                # the regex skips `inf` frames (App. C2)
    assert float(row["PSNR"]) == float("%0.2f" % first)                      # first frame, as ffmpeg prints it
    assert float(row["SSIM"]) == pytest.approx(float("%f" % fr["ssim_all"][0]), abs=2e-6)
    motion, dct, hist, edge, orb, color, tdct, fps = vals
    # the reference's positional unpack (video_processing.py:235-242): columns carry the values in RETURN order
    want = dict(zip(vp.CSV_COLUMNS[7:], (motion, dct, hist, edge, orb, color, tdct, fps)))
    for name, v in want.items():
        assert float(row[name]) == pytest.approx(float(v), rel=rtol, abs=1e-9), name
    assert _stats_files() == cli["stats_before"]                             # temp logs removed (:263-267)


def test_cli_flow_with_oracle_kernels(vqa, cli, monkeypatch):
    from rtvqa_b200 import _native as N
    from rtvqa_b200 import video_processing as vp
    fr, vals, enc_frames = _expected(cli)

    def fake_fr(main_planes, ref_planes, device=None):
        got = RP.psnr_ssim_frames(tuple(main_planes), tuple(ref_planes))
        rows = np.zeros(len(got["psnr_avg"]), dtype=N.FR_DTYPE)
        with np.errstate(divide="ignore"):
            rows["psnr"] = 10.0 * np.log10(255.0 * 255.0 / got["mse"])
        for k in ("mse", "mse_avg", "psnr_avg", "ssim", "ssim_all"):
            rows[k] = got[k]
        return rows

    def fake_complexity(video_path, rw, rh, frame_interval=10, **kw):
        frames = cli["tool"]["decode"](video_path)
        assert video_path.endswith("encoded_video.mp4")                      # complexity of the ENCODED file (:242)
        return RP.average_scene_complexity(np.stack(frames), rw, rh, frame_interval=frame_interval)

    monkeypatch.setattr(vp, "psnr_ssim_frames", fake_fr)
    monkeypatch.setattr(vp, "FR_CHUNK_FRAMES", 4)                            # several bounded reads of the two decode pipes
    chunks = list(vp._yuv420_chunks(cli["src"], 128, 96, chunk=4))
    assert [len(c[0]) for c in chunks] == [4] * (len(enc_frames) // 4) + ([len(enc_frames) % 4] if len(enc_frames) % 4 else [])
    monkeypatch.setattr(vp, "calculate_average_scene_complexity", fake_complexity)
    cfg_file = cli["tmp"] / "config.json"
    cfg_file.write_text(__import__("json").dumps(CFG))
    monkeypatch.setattr("sys.argv", ["video_processing.py", str(cfg_file), cli["src"]])
    vp.main()                                                                # argparse -> load_config -> pipeline
    _check_csv(cli, fr, vals, rtol=1e-12)
    with pytest.raises(FileNotFoundError):
        vp.process_video_and_extract_metrics(str(cli["tmp"] / "missing.mp4"), CFG)
    # anything but 8-bit 4:2:0 is refused rather than scored after a silent conversion
    monkeypatch.setenv("FAKE_FFPROBE_PIX_FMT", "yuv444p10le")
    with pytest.raises(NotImplementedError):
        vp.run_ffmpeg_metrics(cli["src"], cli["src"], str(cli["tmp"] / "p.log"), str(cli["tmp"] / "s.log"), str(cli["tmp"] / "v.json"))


@pytest.mark.gpu
def test_cli_flow_on_device(vqa, cli):
    from rtvqa_b200 import video_processing as vp
    fr, vals, _ = _expected(cli)
    vp.process_video_and_extract_metrics(cli["src"], CFG)
    _check_csv(cli, fr, vals, rtol=1e-4)
