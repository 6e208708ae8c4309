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
