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
