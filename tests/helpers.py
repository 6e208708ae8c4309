"""Shared test doubles."""


class FakeCapture:
    """cv2.VideoCapture stand-in over an in-memory clip: read() / get(CAP_PROP_POS_MSEC) / isOpened() /
    release() with OpenCV's behaviour (POS_MSEC is the timestamp of the frame just read)."""

    def __init__(self, frames, fps=30.0, opened=True, fail_at=None):
        self.frames, self.fps, self.k, self.opened, self.fail_at = frames, float(fps), -1, opened, fail_at
        self.released = False

    def isOpened(self):
        return self.opened and not self.released

    def read(self):
        if self.fail_at is not None and self.k + 1 == self.fail_at:
            raise RuntimeError("decoder blew up")
        if self.k + 1 >= len(self.frames):
            return False, None
        self.k += 1
        return True, self.frames[self.k]

    def get(self, prop):
        assert prop == 0                      # CAP_PROP_POS_MSEC
        return 1000.0 * max(self.k, 0) / self.fps

    def release(self):
        self.released = True


class SparseClip:
    """A clip of ``n`` frames of which only some are materialised: enough for the reference's sampling
    (``len(clip)`` and ``clip[i]`` on sampled indices) without holding 300 x 1080p in memory."""

    def __init__(self, n, frames):
        self.n, self.frames = n, dict(frames)

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return self.frames[i]


_C1_CACHE = {}


def config1_sampled_frames(synth, n=300, h=1080, w=1920, interval=10, seed=0):
    """Sampled frames s_j = source index (j+1)*I - 1 of BASELINE.json config 1 (the reference's own
    CPU-runnable case: 300 x 1080p, frame_interval 10): {source index: frame}, 30 frames."""
    key = (n, h, w, interval, seed)
    if key not in _C1_CACHE:
        _C1_CACHE[key] = {i: f for i, f in enumerate(synth.synth_frame_iter(n, h, w, seed)) if (i + 1) % interval == 0}
    return _C1_CACHE[key]


# ---------------------------------------------------------------------------------------------------------
# Stand-ins for the ffmpeg / ffprobe EXECUTABLES (the image has neither): enough of their command lines for
# video_processing.process_video_and_extract_metrics (reference video_processing.py:180-267) to run end to end.
FAKE_TOOL = r'''#!%(python)s
"""Test double of the ffmpeg / ffprobe command lines the reference issues (tests/helpers.py)."""
import json, re, sys
import os
import cv2
import numpy as np

FAKE_PIX_FMT = os.environ.get("FAKE_FFPROBE_PIX_FMT", "yuv420p")


def decode(path):
    cap, out = cv2.VideoCapture(path), []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        out.append(f)
    cap.release()
    return out


def encode(src, dst):
    """'libx264 -crf' stand-in: a mild blur, re-encoded with the mp4v writer OpenCV ships."""
    frames = decode(src)
    h, w = frames[0].shape[:2]
    wr = cv2.VideoWriter(dst, cv2.VideoWriter_fourcc(*"mp4v"), 30.0, (w, h))
    assert wr.isOpened()
    for f in frames:
        wr.write(cv2.GaussianBlur(f, (3, 3), 0))
    wr.release()


def main(argv):
    tool = argv[0].rsplit("/", 1)[-1]
    a = argv[1:]
    if tool == "ffprobe":
        f = decode(a[-1])[0]
        print(json.dumps({"streams": [{"width": f.shape[1], "height": f.shape[0], "avg_frame_rate": "30/1", "bit_rate": "1234567",
                                       "pix_fmt": FAKE_PIX_FMT}]}))
    elif "-c:v" in a:
        encode(a[a.index("-i") + 1], a[-1])
    elif "rawvideo" in a:
        for f in decode(a[a.index("-i") + 1]):
            sys.stdout.buffer.write(cv2.cvtColor(f, cv2.COLOR_BGR2YUV_I420).tobytes())
    elif "-filter_complex" in a:
        log = re.search(r"log_path=([^:]+)", a[a.index("-filter_complex") + 1]).group(1)
        with open(log, "w") as fh:
            json.dump({"pooled_metrics": {"vmaf": {"mean": 93.25}}}, fh)
    else:
        sys.exit("fake ffmpeg: unsupported command line: %%r" %% (a,))


if __name__ == "__main__":
    main(sys.argv)
'''


def install_fake_ffmpeg(bin_dir):
    """Write executable `ffmpeg` and `ffprobe` doubles into bin_dir; returns the module namespace of the
    double (decode / encode) for computing expectations."""
    import os
    import sys
    src = FAKE_TOOL % {"python": sys.executable}
    for name in ("ffmpeg", "ffprobe"):
        p = os.path.join(str(bin_dir), name)
        with open(p, "w") as f:
            f.write(src)
        os.chmod(p, 0o755)
    ns = {"__name__": "fake_ffmpeg"}
    exec(compile(src, "fake_ffmpeg", "exec"), ns)
    return ns


def yuv420_planes(frames):
    """[n,h,w] Y and [n,h/2,w/2] U, V of BGR frames, the way the ffmpeg double's rawvideo output is laid out."""
    import cv2
    import numpy as np
    h, w = frames[0].shape[:2]
    a = np.stack([cv2.cvtColor(f, cv2.COLOR_BGR2YUV_I420).reshape(-1) for f in frames])
    y = a[:, :h * w].reshape(-1, h, w)
    u = a[:, h * w:h * w * 5 // 4].reshape(-1, h // 2, w // 2)
    v = a[:, h * w * 5 // 4:].reshape(-1, h // 2, w // 2)
    return np.ascontiguousarray(y), np.ascontiguousarray(u), np.ascontiguousarray(v)
