"""Host-side logic of the asynchronous sample pipeline (mvster_b200/prefetch.py) on a CPU device: order, bounded read-ahead,
error position, nested structures, and the evaluation jobs on a synthetic scan in the reference's directory layout.  The CUDA
side (pinned staging, copy stream, events) is covered by tests/test_gpu_zzzz_prefetch.py."""
import threading
import time

import numpy as np
import pytest
import torch

from test_formats import _write_scan
from mvster_b200 import formats, prefetch


def test_order_structure_and_passthrough_leaves():
    def job(i):
        def f():
            time.sleep(0.01 * ((7 * i) % 4))            # jobs finish out of order
            return {"imgs": [np.full((2, 3), i, np.float32), torch.full((1,), float(i))], "meta": ("name%d" % i, i),
                    "nested": {"k": (np.arange(i + 1),)}}
        return f
    got = list(prefetch.Prefetcher([job(i) for i in range(9)], device="cpu", depth=2, workers=3))
    assert len(got) == 9
    for i, s in enumerate(got):
        assert isinstance(s["imgs"], list) and isinstance(s["imgs"][0], torch.Tensor) and s["imgs"][0].dtype == torch.float32
        assert torch.equal(s["imgs"][0], torch.full((2, 3), float(i))) and s["imgs"][1].item() == i
        assert s["meta"] == ("name%d" % i, i)
        assert isinstance(s["nested"]["k"], tuple) and s["nested"]["k"][0].tolist() == list(range(i + 1))


def test_read_ahead_is_bounded_and_lazy():
    started, lock = [], threading.Lock()

    def jobs():
        for i in range(50):
            def f(i=i):
                with lock:
                    started.append(i)
                return {"x": np.zeros(1, np.float32) + i}
            yield f
    it = iter(prefetch.Prefetcher(jobs(), device="cpu", depth=2, workers=2))
    first = next(it)
    time.sleep(0.2)
    assert first["x"].item() == 0
    assert len(started) <= 1 + 2 * (2 + 2)              # nothing like the whole list is decoded ahead of the consumer
    assert [s["x"].item() for s in it] == list(range(1, 50))
    assert sorted(started) == list(range(50))


def test_a_failing_job_raises_at_its_position():
    def ok(i):
        return lambda: {"i": np.array([i])}

    def bad():
        raise OSError("unreadable view")
    seen = []
    with pytest.raises(OSError, match="unreadable view"):
        for s in prefetch.Prefetcher([ok(0), ok(1), bad, ok(3)], device="cpu", depth=1, workers=2):
            seen.append(int(s["i"][0]))
    assert seen == [0, 1]
    with pytest.raises(ValueError):
        prefetch.Prefetcher([], depth=0)


def test_transform_and_eval_jobs_on_a_synthetic_scan(tmp_path):
    _write_scan(tmp_path)
    pairs = formats.read_pair_file(str(tmp_path / "scan1" / "pair.txt"), 4)
    jobs = prefetch.eval_jobs(str(tmp_path), "scan1", pairs, 4, interval_scale=1.06)
    assert len(jobs) == len(pairs) == 3                 # the view without sources is dropped, like the reference does
    seen = 0
    for (ref, srcs), s in zip(pairs, prefetch.Prefetcher(jobs, device="cpu", transform=lambda s: {**s, "tag": 1})):
        want = formats.load_eval_sample(str(tmp_path), "scan1", ref, srcs, 4, 1.06)
        assert s["tag"] == 1 and s["filename"] == [want["filename"]]
        assert len(s["imgs"]) == 4 and all(i.shape == (1, 3, 128, 192) for i in s["imgs"])
        assert all(torch.equal(a[0], torch.from_numpy(b)) for a, b in zip(s["imgs"], want["imgs"]))
        assert s["depth_values"].shape == (1, 192)
        for k in ("stage1", "stage2", "stage3", "stage4"):
            assert s["proj_matrices"][k].shape == (1, 4, 2, 4, 4)
            assert np.array_equal(s["proj_matrices"][k][0].numpy(), want["proj_matrices"][k])
        seen += 1
    assert seen == 3


def test_consumer_can_stop_early_without_hanging():
    started = []

    def jobs():
        for i in range(1000):
            def f(i=i):
                started.append(i)
                time.sleep(0.005)
                return {"x": np.zeros(2, np.float32) + i}
            yield f
    t0 = time.time()
    for s in prefetch.Prefetcher(jobs(), device="cpu", depth=2, workers=2):
        if s["x"][0].item() == 3:
            break
    assert time.time() - t0 < 5.0 and len(started) < 40       # pending jobs are cancelled, the pool is shut down
