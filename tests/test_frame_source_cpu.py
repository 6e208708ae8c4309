"""SampledFrameSource: one decode pass must reproduce the sampling of the reference's
read_frame_pairs (complexity_metrics.py:76-111) and extract_frame_timestamps (:38-73) exactly."""
import os
import sys

import numpy as np
import pytest

from helpers import FakeCapture
from oracle import np_oracle as NO


def _clip(n, h=48, w=64, seed=0):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)


def _source(vqa, frames, interval, chunk, **kw):
    from rtvqa_b200.frame_source import SampledFrameSource
    return SampledFrameSource("clip.mp4", interval, chunk, capture_factory=lambda p: FakeCapture(frames, 30.0, **kw))


@pytest.mark.parametrize("n,interval,chunk", [(47, 1, 8), (47, 3, 4), (47, 10, 48), (30, 10, 1), (9, 10, 4), (10, 10, 4), (0, 5, 4)])
def test_sampling_matches_reference_rules(vqa, n, interval, chunk):
    frames = _clip(n)
    src = _source(vqa, frames, interval, chunk)
    chunks = list(src)
    want_idx = NO.sampled_indices(n, interval)                   # I-1, 2I-1, ...
    got = np.concatenate(chunks) if chunks else np.empty((0,) + frames.shape[1:], np.uint8)
    assert got.shape[0] == len(want_idx)
    assert np.array_equal(got, frames[want_idx]) if len(want_idx) else True
    assert all(c.shape[0] == chunk for c in chunks[:-1]) and all(c.flags["C_CONTIGUOUS"] for c in chunks)
    assert chunks[-1].shape[0] <= chunk if chunks else True
    want_ts = [1000.0 * i / 30.0 for i in NO.timestamp_indices(n, interval)]   # 0, I, 2I, ...
    assert src.timestamps == want_ts
    assert src.frames_decoded == n and src.frames_sampled == len(want_idx) and src.opened is True


def test_unopenable_file_yields_nothing(vqa):
    src = _source(vqa, _clip(5), 1, 4, opened=False)
    assert list(src) == [] and src.timestamps == [] and src.opened is False


def test_decoder_error_is_raised_in_the_consumer(vqa):
    src = _source(vqa, _clip(20), 1, 4, fail_at=9)
    got = []
    with pytest.raises(RuntimeError, match="decoder blew up"):
        for c in src:
            got.append(c)
    assert sum(len(c) for c in got) == 8                         # two full chunks arrived before the failure


def test_early_exit_stops_the_decode_thread(vqa):
    src = _source(vqa, _clip(200), 1, 4)
    for i, c in enumerate(src):
        if i == 1:
            break
    src.close()
    assert not src._thread.is_alive()
    with pytest.raises(RuntimeError):
        iter(src).__next__()


def test_bad_arguments(vqa):
    from rtvqa_b200.frame_source import SampledFrameSource
    with pytest.raises(ValueError):
        SampledFrameSource("a.mp4", 0)
    with pytest.raises(ValueError):
        SampledFrameSource("a.mp4", 1, 0)


def test_real_container_against_the_module_readers(vqa, tmp_path):
    """An actual file through cv2: the streaming source equals the two reference-shaped readers
    (three decodes in the reference, one here)."""
    cv2 = pytest.importorskip("cv2")
    from rtvqa_b200 import complexity_metrics as cm
    from rtvqa_b200.frame_source import SampledFrameSource
    path = str(tmp_path / "clip.avi")
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 25.0, (96, 64))
    if not wr.isOpened():
        pytest.skip("no MJPG writer in this OpenCV build")
    frames = _clip(23, 64, 96, seed=3)
    for f in frames:
        wr.write(f)
    wr.release()
    for interval in (1, 4, 10):
        pairs = cm.read_frame_pairs(path, interval)
        stamps = cm.extract_frame_timestamps(path, interval)
        src = SampledFrameSource(path, interval, 5)
        chunks = list(src)
        got = np.concatenate(chunks) if chunks else None
        if pairs:
            want = np.stack([pairs[0][1]] + [p[0] for p in pairs])
            assert np.array_equal(got, want)
        else:
            assert got is None or len(got) <= 1
        assert src.timestamps == stamps
