"""`nn.DataParallel(model)` as `test_mvs4.py:196` wraps it: one host thread per device runs `MVS4net.forward` concurrently.
The library keeps no process-global mutable state on that path (the SM cap of the persistent convolution kernels is
thread-local, the SM count is per device) and CUDA-graph capture uses the thread-local error mode, so replicas must reproduce
the single-device result bit for bit - also when both capture their graphs at the same time.  Needs two GPUs
(`gpurun --gpus 2`); on one GPU the same check runs with two threads sharing the device through two model instances."""
import threading

import pytest
import torch

from util import SHIPPED, build_model
from mvster_b200 import synth

pytestmark = pytest.mark.gpu


def _to(dev, imgs, proj, dv):
    return [t.to(dev) for t in imgs], {k: v.to(dev) for k, v in proj.items()}, dv.to(dev)


@pytest.mark.parametrize("graph", [False, True], ids=["eager", "graph"])
def test_two_host_threads_reproduce_the_single_thread_result(graph):
    ndev = torch.cuda.device_count()
    devs = [torch.device("cuda", i % ndev) for i in range(2)]
    frames = [synth.make_inputs(1, 3, 128, 192, seed=40 + i) for i in range(2)]
    models = [build_model(SHIPPED, 3).to(d) for d in devs]
    for m in models:
        m.use_cuda_graph = graph
    want = []
    with torch.no_grad():
        for m, d, f in zip(models, devs, frames):
            out = m(*_to(d, *f))
            want.append({k: out[k].clone() for k in ("depth", "photometric_confidence", "attn_weight")})
    torch.cuda.synchronize()
    if graph:  # make the threads capture (not just replay) concurrently
        for m in models:
            for eng in m._engines.values():
                eng._graphs.clear()
    got, errs = [None, None], []
    gate = threading.Barrier(2)

    def worker(i):
        try:
            with torch.no_grad():
                torch.cuda.set_device(devs[i])
                gate.wait()
                for _ in range(3):
                    out = models[i](*_to(devs[i], *frames[i]))
                torch.cuda.synchronize(devs[i])
                got[i] = out
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for i in range(2):
        for k, v in want[i].items():
            assert torch.equal(got[i][k], v), (i, k)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_nn_dataparallel_wrapper_matches_single_device():
    imgs, proj, dv = synth.make_inputs(2, 3, 128, 192, seed=44)
    m = build_model(SHIPPED, 3).to("cuda:0")
    with torch.no_grad():
        want = m(*_to("cuda:0", imgs, proj, dv))["depth"].clone()
        dp = torch.nn.DataParallel(m, device_ids=[0, 1])
        got = dp(*_to("cuda:0", imgs, proj, dv))["depth"]
    assert torch.equal(got.cpu(), want.cpu())
