"""Host mirror of the reference's ``video_processing.py`` (same names, config knobs, CSV
columns and error behaviour).  PSNR and SSIM -- which the reference delegates to FFmpeg's
``psnr``/``ssim`` filters in ``run_ffmpeg_metrics`` (video_processing.py:270-297) -- run in the
fused sm_100a kernel of libvqa_b200.so; VMAF stays on FFmpeg/libvmaf; the libx264 encode, the
ffprobe call and container decode stay on the FFmpeg executables (out of scope).

Unlike the reference this module has no import-time side effects (the reference opens
``video_processing.log`` and starts a QueueListener at import, video_processing.py:26-38).
"""
from __future__ import annotations

import argparse
import csv
import json
import logging
import multiprocessing
import os
import re
import shutil
import subprocess
import tempfile
import threading
import uuid
from fractions import Fraction

import numpy as np

try:
    from . import _native as N
    from .complexity_metrics import calculate_average_scene_complexity
except ImportError:  # top-level drop-in use
    import _native as N  # type: ignore
    from complexity_metrics import calculate_average_scene_complexity  # type: ignore

logging.basicConfig(level=logging.INFO)
logger = logging.getLogger(__name__)
log_lock = threading.Lock()

CSV_COLUMNS = ['Bitrate (kbps)', 'Resolution (px)', 'Frame Rate (fps)', 'CRF', 'PSNR', 'SSIM', 'VMAF',
               'Advanced Motion Complexity', 'DCT Complexity', 'Temporal DCT Complexity', 'Histogram Complexity',
               'Edge Detection Complexity', 'ORB Feature Complexity', 'Color Histogram Complexity',
               'Framerate Variation']


def thread_safe_update_csv(metrics, csv_file='video_quality_data.csv'):
    """Append one row; header iff the file did not exist (reference :44-68).  Column order is the
    dict's insertion order, as with ``pd.DataFrame([metrics]).to_csv(index=False)``."""
    file_exists = os.path.isfile(csv_file)
    with log_lock:
        try:
            with open(csv_file, 'a', newline='') as f:
                w = csv.writer(f, lineterminator='\n')
                if not file_exists:
                    w.writerow(list(metrics.keys()))
                # DataFrame.to_csv writes missing values (None / NaN) as empty fields
                w.writerow(["" if v is None or (isinstance(v, (float, np.floating)) and np.isnan(v))
                            else (repr(float(v)) if isinstance(v, (float, np.floating)) else v)
                            for v in metrics.values()])
        except IOError as e:
            logger.error("Failed to write to CSV file: %s", e)
            raise


def load_config(config_file):
    """Load + validate the JSON config (reference :71-85)."""
    try:
        with open(config_file, 'r') as f:
            config = json.load(f)
        validate_config(config)
        return config
    except FileNotFoundError:
        logger.error(f"Configuration file {config_file} not found.")
        raise
    except json.JSONDecodeError:
        logger.error(f"Error decoding JSON from configuration file {config_file}.")
        raise


def validate_config(config):
    """Same checks and messages as the reference (:87-98)."""
    if not (1 <= config.get('crf', 23) <= 51):
        raise ValueError("CRF value must be between 1 and 51.")
    if config.get('resize_width', 0) <= 0 or config.get('resize_height', 0) <= 0:
        raise ValueError("Resize dimensions must be positive integers.")
    if config.get('frame_interval', 10) <= 0:
        raise ValueError("Frame interval must be a positive integer.")
    if not isinstance(config.get('num_workers', multiprocessing.cpu_count() // 2), int):
        raise ValueError("num_workers must be an integer.")
    # optional extension: ORB input size; absent = the reference's hard-wired 64x64
    ow, oh = config.get('orb_width'), config.get('orb_height')
    if (ow is None) != (oh is None) or (ow is not None and (not isinstance(ow, int) or not isinstance(oh, int)
                                                            or ow <= 0 or oh <= 0)):
        raise ValueError("orb_width and orb_height must both be positive integers (or both absent).")


def get_video_info(video_path):
    """ffprobe -> (bitrate kbps, 'WxH', frame rate, width, height) (reference :101-142).  The
    reference eval()s avg_frame_rate; a Fraction parse gives the same number safely."""
    cmd = ['ffprobe', '-v', 'error', '-select_streams', 'v:0', '-print_format', 'json', '-show_entries',
           'stream=width,height,avg_frame_rate,bit_rate', video_path]
    try:
        process = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        stdout, _ = process.communicate()
        if process.returncode != 0:
            logger.error("ffprobe failed to retrieve video information.")
            raise RuntimeError("ffprobe failed.")
        stream = json.loads(stdout)['streams'][0]
        bitrate = int(stream.get('bit_rate', 0)) // 1000
        width, height = stream.get('width', 0), stream.get('height', 0)
        rate = stream.get('avg_frame_rate', '0/1')
        frame_rate = float(Fraction(rate)) if rate != '0/0' else 0
        return bitrate, f"{width}x{height}", frame_rate, width, height
    except Exception as e:
        logger.error(f"Error retrieving video information: {e}")
        raise


def extract_metrics_from_logs(psnr_log, ssim_log, vmaf_log, video_file, crf, bitrate, resolution, frame_rate):
    """First ``psnr_avg:`` and first ``All:`` of the per-frame stats files, pooled VMAF mean
    (reference :145-177 -- the CSV carries first-frame PSNR/SSIM)."""
    metrics = {'Bitrate (kbps)': bitrate, 'Resolution (px)': resolution, 'Frame Rate (fps)': frame_rate, 'CRF': crf}
    try:
        if os.path.isfile(psnr_log):
            with open(psnr_log) as f:
                m = re.search(r'psnr_avg:(\s*\d+\.\d+)', f.read())
                if m:
                    metrics['PSNR'] = float(m.group(1))
        if os.path.isfile(ssim_log):
            with open(ssim_log) as f:
                m = re.search(r'All:(\s*\d+\.\d+)', f.read())
                if m:
                    metrics['SSIM'] = float(m.group(1))
        if os.path.isfile(vmaf_log):
            with open(vmaf_log) as f:
                data = json.load(f)
                if 'pooled_metrics' in data and 'vmaf' in data['pooled_metrics']:
                    metrics['VMAF'] = data['pooled_metrics']['vmaf']['mean']
    except Exception as e:
        logger.warning(f"Error extracting metrics from logs: {e}")
    return metrics


# --------------------------------------------------------------------------- PSNR / SSIM on the GPU
def psnr_ssim_frames(main_planes, ref_planes, device=None):
    """Per-frame PSNR/SSIM of planar 8-bit stacks ((Y,U,V), each [n,h_c,w_c]); ``main`` is the
    distorted input [0:v], ``ref`` the reference [1:v] as in the reference's filter graph.
    Returns the structured array of ``_native.FR_DTYPE``."""
    return N.get_context(device, role="fr").psnr_ssim(main_planes, ref_planes)


def analyze_frames(frames, resize_width, resize_height, main_planes, ref_planes, device=None, orb_size=None):
    """Both halves of ``process_video_and_extract_metrics`` (reference :216 PSNR/SSIM, :242 scene
    complexity) for pre-decoded HOST buffers in one device pass with one interleaved upload schedule
    (``vqa_analyze_clip``): ``frames`` (n,h,w,3) uint8 BGR sampled frames, planes 3 x [n_pairs,h_c,w_c]
    uint8 (distorted = main, reference = ref).  Returns (complexity rows, PSNR/SSIM rows); identical to
    ``complexity_metrics._clip_metrics`` + ``psnr_ssim_frames`` called in turn."""
    return N.get_context(device).analyze_clip(frames, resize_width, resize_height, main_planes, ref_planes,
                                              orb_size=orb_size)


def _fmt_psnr(v):
    return "inf" if np.isinf(v) else "%0.2f" % v


def write_ffmpeg_stats(rows, psnr_log, ssim_log):
    """Write per-frame stats files in the layout FFmpeg's psnr/ssim filters print, so that
    ``extract_metrics_from_logs`` (and any other parser of those files) keeps working."""
    with open(psnr_log, 'w') as f:
        for i, r in enumerate(rows, 1):
            f.write("n:%d mse_avg:%0.2f mse_y:%0.2f mse_u:%0.2f mse_v:%0.2f psnr_avg:%s psnr_y:%s psnr_u:%s psnr_v:%s \n" % (
                i, r['mse_avg'], r['mse'][0], r['mse'][1], r['mse'][2], _fmt_psnr(r['psnr_avg']),
                _fmt_psnr(r['psnr'][0]), _fmt_psnr(r['psnr'][1]), _fmt_psnr(r['psnr'][2])))
    with open(ssim_log, 'w') as f:
        for i, r in enumerate(rows, 1):
            a = float(r['ssim_all'])
            db = float('inf') if a >= 1.0 else 10.0 * np.log10(1.0 / (1.0 - a))
            f.write("n:%d Y:%f U:%f V:%f All:%f (%f)\n" % (i, r['ssim'][0], r['ssim'][1], r['ssim'][2], a, db))


def _probe_size(path):
    _, _, _, w, h = get_video_info(path)
    return int(w), int(h)


FR_CHUNK_FRAMES = 64          # frames per PSNR/SSIM device call: host memory stays O(chunk), like ffmpeg's own O(1) filters
_YUV420_8BIT = ('yuv420p', 'yuvj420p', 'nv12', 'nv21')


def _probe_pix_fmt(path):
    """pix_fmt of the first video stream, or None when ffprobe does not report one."""
    cmd = ['ffprobe', '-v', 'error', '-select_streams', 'v:0', '-print_format', 'json', '-show_entries',
           'stream=pix_fmt', path]
    try:
        out = subprocess.run(cmd, check=True, stdout=subprocess.PIPE).stdout
        return json.loads(out)['streams'][0].get('pix_fmt')
    except Exception:
        return None


def _yuv420_chunks(path, w, h, chunk):
    """Decode to planar yuv420p with the ffmpeg executable (container decode is out of scope) and yield
    (Y [n,h,w], U, V [n,ceil(h/2),ceil(w/2)]) stacks of at most ``chunk`` frames read from the pipe as they come."""
    cw, ch = (w + 1) // 2, (h + 1) // 2
    ysz, csz = w * h, cw * ch
    fsz = ysz + 2 * csz
    cmd = ['ffmpeg', '-v', 'error', '-i', path, '-f', 'rawvideo', '-pix_fmt', 'yuv420p', '-']
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE)
    try:
        while True:
            parts, need = [], fsz * chunk
            while need > 0:
                b = proc.stdout.read(need)
                if not b:
                    break
                parts.append(b)
                need -= len(b)
            raw = b''.join(parts)
            n = len(raw) // fsz
            if n == 0:
                break
            a = np.frombuffer(raw, np.uint8, n * fsz).reshape(n, fsz)
            yield (np.ascontiguousarray(a[:, :ysz]).reshape(n, h, w),
                   np.ascontiguousarray(a[:, ysz:ysz + csz]).reshape(n, ch, cw),
                   np.ascontiguousarray(a[:, ysz + csz:]).reshape(n, ch, cw))
            if need > 0:
                break
    finally:
        proc.stdout.close()
        rc = proc.wait()
    if rc != 0:
        raise subprocess.CalledProcessError(rc, cmd)


def run_ffmpeg_metrics(reference_video, distorted_video, psnr_log, ssim_log, vmaf_log, vmaf_model_path=None):
    """PSNR + SSIM on the GPU, VMAF through FFmpeg/libvmaf (reference :270-297).  Produces the
    same three files the reference's single ffmpeg invocation produces.  The two decodes are read from
    their pipes in bounded chunks (``FR_CHUNK_FRAMES``) and scored chunk by chunk.  FFmpeg's psnr/ssim
    filters work in the main input's own pixel format; this path covers 8-bit 4:2:0 and says so for
    anything else instead of scoring a silently converted picture.  Decode failures and a failing
    libvmaf run are logged and re-raised like the reference's CalledProcessError path."""
    try:
        w, h = _probe_size(distorted_video)
        fmt = _probe_pix_fmt(distorted_video)
        if fmt is not None and fmt not in _YUV420_8BIT:
            raise NotImplementedError(f"PSNR/SSIM on the device covers 8-bit 4:2:0 video; {distorted_video} is {fmt}")
        parts = []
        for main, ref in zip(_yuv420_chunks(distorted_video, w, h, FR_CHUNK_FRAMES),
                             _yuv420_chunks(reference_video, w, h, FR_CHUNK_FRAMES)):
            n = min(len(main[0]), len(ref[0]))
            parts.append(psnr_ssim_frames([p[:n] for p in main], [p[:n] for p in ref]))
            if len(main[0]) != len(ref[0]):
                break
        rows = np.concatenate(parts) if parts else np.zeros(0, dtype=N.FR_DTYPE)
        write_ffmpeg_stats(rows, psnr_log, ssim_log)
        if vmaf_model_path and os.path.isfile(vmaf_model_path):
            flt = f"[0:v][1:v]libvmaf=model_path={vmaf_model_path}:log_path={vmaf_log}:log_fmt=json"
        else:
            flt = f"[0:v][1:v]libvmaf=log_path={vmaf_log}:log_fmt=json"
        subprocess.run(['ffmpeg', '-i', distorted_video, '-i', reference_video, '-filter_complex', flt,
                        '-f', 'null', '-'], check=True)
    except (subprocess.CalledProcessError, FileNotFoundError) as e:
        logger.error(f"FFmpeg metrics calculation failed: {e}")
        raise


def process_video_and_extract_metrics(input_video, config, correct_column_order=False):
    """Encode (libx264 CRF) -> PSNR/SSIM/VMAF -> scene complexity of the ENCODED file -> CSV row
    (reference :180-267).  By default the reference's positional unpack is reproduced, i.e. the
    CSV columns carry the values the reference writes under them (the
    returned order is motion, dct, hist, edge, orb, colour, temporal-dct, fps but it is unpacked
    as motion, dct, temporal-dct, hist, edge, orb, colour, fps).  ``correct_column_order=True``
    writes each value under its true name instead."""
    crf = config.get("crf", 23)
    vmaf_model_path = config.get("vmaf_model_path", None)
    resize_width = config.get("resize_width", 64)
    resize_height = config.get("resize_height", 64)
    frame_interval = config.get("frame_interval", 10)
    uid = uuid.uuid4().hex
    psnr_log = os.path.join(tempfile.gettempdir(), f'psnr_{uid}.log')
    ssim_log = os.path.join(tempfile.gettempdir(), f'ssim_{uid}.log')
    vmaf_log = os.path.join(tempfile.gettempdir(), f'vmaf_{uid}.json')
    if not os.path.isfile(input_video):
        raise FileNotFoundError(f"The input video file {input_video} does not exist.")
    temp_dir = tempfile.mkdtemp()
    try:
        encoded_video = os.path.join(temp_dir, 'encoded_video.mp4')
        try:
            subprocess.run(['ffmpeg', '-i', input_video, '-c:v', 'libx264', '-crf', str(crf), '-preset', 'medium',
                            '-y', encoded_video], check=True)
        except subprocess.CalledProcessError as e:
            logger.error(f"FFmpeg encoding failed: {e}")
            raise
        run_ffmpeg_metrics(input_video, encoded_video, psnr_log, ssim_log, vmaf_log, vmaf_model_path)
        bitrate, resolution, frame_rate, _, _ = get_video_info(input_video)
        metrics = extract_metrics_from_logs(psnr_log=psnr_log, ssim_log=ssim_log, vmaf_log=vmaf_log,
                                            video_file=input_video, crf=crf, bitrate=bitrate,
                                            resolution=resolution, frame_rate=frame_rate)
        logger.info("Metrics extracted: %s", metrics)
        logger.info("Calculating scene complexity after encoding...")
        orb_size = (config['orb_width'], config['orb_height']) if config.get('orb_width') else None
        vals = calculate_average_scene_complexity(encoded_video, resize_width, resize_height,
                                                  frame_interval=frame_interval, orb_size=orb_size)
        metrics.update(complexity_columns(vals, correct_column_order))
        thread_safe_update_csv(metrics, csv_file='video_quality_data.csv')
    finally:
        for log_file in (psnr_log, ssim_log, vmaf_log):
            if os.path.exists(log_file):
                os.remove(log_file)
        shutil.rmtree(temp_dir)


def complexity_columns(vals, correct_column_order=False):
    """Map the 8-tuple of calculate_average_scene_complexity onto the CSV columns."""
    motion, dct, hist, edge, orb, color, tdct, fps = vals
    if correct_column_order:
        a = (motion, dct, tdct, hist, edge, orb, color, fps)
    else:   # the reference's positional unpack (video_processing.py:235-242)
        a = (motion, dct, hist, edge, orb, color, tdct, fps)
    names = CSV_COLUMNS[7:]
    return dict(zip(names, a))


def main():
    parser = argparse.ArgumentParser(description="Process a video, extract metrics, and update CSV.")
    parser.add_argument('config_file', type=str, help="Path to the configuration JSON file.")
    parser.add_argument('input_video', type=str, help="Path to the input video file.")
    args = parser.parse_args()
    config = load_config(args.config_file)
    try:
        process_video_and_extract_metrics(input_video=args.input_video, config=config)
        logger.info("Processing completed successfully.")
    except Exception as e:
        logger.error(f"An error occurred during processing: {e}")
        raise


if __name__ == "__main__":
    main()
