"""The asynchronous sample pipeline on the GPU (pinned staging ring, copy stream, events) and the whole evaluation flow it
feeds on a synthetic scan in the reference's directory layout: read -> forward -> PFM depth / confidence maps ->
geometric-consistency filter -> PLY, the sequence of test_mvs4.py:199-262 and :332-421."""
import numpy as np
import pytest
import torch

from flow_util import scan_to_point_cloud
from util import SHIPPED, build_model

from mvster_b200 import prefetch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_samples_arrive_intact_in_order_and_buffers_are_reused():
    rng = np.random.RandomState(0)
    data = [{"a": rng.rand(3, 64, 96).astype(np.float32), "b": [rng.rand(2, 4, 4).astype(np.float32)], "i": i} for i in range(12)]
    pf = prefetch.Prefetcher([(lambda d=d: d) for d in data], device=DEV, depth=2, workers=3)
    acc = []
    for s in pf:
        assert s["a"].is_cuda and s["b"][0].is_cuda
        acc.append((s["i"], s["a"] * 2.0, s["b"][0].clone()))      # consumed on the current stream, after the copy event
    torch.cuda.synchronize()
    assert [i for i, _, _ in acc] == list(range(12))
    for (i, a2, b), d in zip(acc, data):
        assert torch.equal(a2.cpu(), torch.from_numpy(d["a"]) * 2.0) and torch.equal(b.cpu(), torch.from_numpy(d["b"][0]))
    assert pf.h2d_bytes == sum(d["a"].nbytes + d["b"][0].nbytes for d in data)


def test_scan_to_point_cloud(tmp_path):
    model = build_model(SHIPPED, seed=2).to(DEV).eval()
    with torch.no_grad():
        scan_to_point_cloud(tmp_path, lambda s: model(s["imgs"], s["proj_matrices"], s["depth_values"]), DEV)
