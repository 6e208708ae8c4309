"""Single-decode streaming frame source.

The reference decodes a clip three times per analysis -- ``read_frame_pairs``
(complexity_metrics.py:76-111) keeps every sampled frame of the clip in RAM, ``calculate_temporal_dct``
(:506-541) calls it again, and ``extract_frame_timestamps`` (:38-73) walks the file a third time.
``SampledFrameSource`` walks the file ONCE, in a background thread (OpenCV's decode releases the GIL),
and hands out

  * the sampled frames ``s_j`` = source index ``(j + 1) * I - 1`` (the counter of ``read_frame_pairs``
    is incremented before its modulo test) in chunks of ``chunk_frames`` C-contiguous ``(m, h, w, 3)``
    uint8 BGR arrays, ready for ``vqa_complexity_frames`` with the previous chunk's last frame as halo;
  * the timestamps of source indices ``0, I, 2I, ...`` (``extract_frame_timestamps`` tests the modulo
    before incrementing), i.e. ``CAP_PROP_POS_MSEC`` right after the frame was read;

so that at most ``queue_depth + 1`` chunks are resident, and the decode of chunk ``k + 1`` overlaps the
GPU work on chunk ``k``.  Decoding itself stays on OpenCV/libav: there is no NVDEC binding in this
image, and the reference's frame bytes are by definition what ``cv2.VideoCapture.read()`` returns.
"""
from __future__ import annotations

import logging
import queue
import threading

import numpy as np

logger = logging.getLogger(__name__)


class SampledFrameSource:
    """Iterate over chunks of sampled frames of ``video_path``; ``timestamps`` is complete once the
    iteration has finished.  ``opened`` is False when the file could not be opened (the reference
    logs an error and returns an empty list, complexity_metrics.py:56-58 / :87-89)."""

    def __init__(self, video_path, frame_interval=10, chunk_frames=48, queue_depth=2, capture_factory=None):
        if frame_interval <= 0:
            raise ValueError("frame_interval must be positive")
        if chunk_frames <= 0:
            raise ValueError("chunk_frames must be positive")
        self.video_path = video_path
        self.frame_interval = int(frame_interval)
        self.chunk_frames = int(chunk_frames)
        self.timestamps = []           # ms, source indices 0, I, 2I, ...
        self.frames_decoded = 0
        self.frames_sampled = 0
        self.opened = None
        self._q = queue.Queue(maxsize=max(1, int(queue_depth)))
        self._err = None
        self._stop = threading.Event()
        self._factory = capture_factory
        self._thread = threading.Thread(target=self._run, name="vqa-decode", daemon=True)
        self._started = False

    # ---------------------------------------------------------------- producer (decode thread)
    def _open(self):
        if self._factory is not None:
            return None, self._factory(self.video_path)
        import cv2
        return cv2, cv2.VideoCapture(self.video_path)

    def _put(self, item):
        while not self._stop.is_set():
            try:
                self._q.put(item, timeout=0.1)
                return True
            except queue.Full:
                continue
        return False

    def _run(self):
        cap = None
        try:
            cv2, cap = self._open()
            pos_msec = 0 if cv2 is None else cv2.CAP_PROP_POS_MSEC
            self.opened = bool(cap.isOpened())
            if not self.opened:
                logger.error(f"Error opening video file: {self.video_path}")
                return
            I, buf, fill, count = self.frame_interval, None, 0, 0
            while cap.isOpened() and not self._stop.is_set():
                ok, frame = cap.read()
                if not ok:
                    break
                if count % I == 0:                       # extract_frame_timestamps: test, then increment
                    self.timestamps.append(cap.get(pos_msec))
                count += 1
                if count % I == 0:                       # read_frame_pairs: increment, then test
                    if buf is None or buf.shape[1:] != frame.shape:
                        if fill:                         # frame size changed mid-stream: flush what we have
                            if not self._put(buf[:fill]):
                                return
                        buf, fill = np.empty((self.chunk_frames,) + frame.shape, dtype=np.uint8), 0
                    buf[fill] = frame
                    fill += 1
                    self.frames_sampled += 1
                    if fill == self.chunk_frames:
                        if not self._put(buf):
                            return
                        buf, fill = None, 0
            self.frames_decoded = count
            if fill:
                self._put(buf[:fill])
        except BaseException as e:          # surfaced in the consumer thread
            self._err = e
        finally:
            if cap is not None:
                cap.release()
            self._put(None)

    # ---------------------------------------------------------------- consumer
    def __iter__(self):
        if self._started:
            raise RuntimeError("SampledFrameSource can be iterated once")
        self._started = True
        self._thread.start()
        try:
            while True:
                item = self._q.get()
                if item is None:
                    break
                yield item
        finally:
            self._stop.set()
            self._thread.join(timeout=30)
        if self._err is not None:
            raise self._err

    def close(self):
        self._stop.set()
        if self._started:
            self._thread.join(timeout=30)
