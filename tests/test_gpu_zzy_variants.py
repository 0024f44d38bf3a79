"""The default-off CUDA-core kernel variants (fpn.cu: gather 2 / 3, merge 2 / 3, four-pixel stem; conv_simt_px2.cu: four-voxel
conv0) against the default kernels on the GPU.  They were written after the round's last GPU session and had only run on the
CPU emulation (tests/test_emu_kernels.py, also under Address- and ThreadSanitizer); this file is the GPU half of that check; it
sorts after the tests that were green on the B200 and before the two files with more GPU-only surface (backward, prefetcher).  The library reads its switches at every launch, so the environment is set per call."""
import numpy as np
import pytest
import torch

from util import SHIPPED, build_model

from mvster_b200 import _lib, capi, fpn_engine, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _t(a):
    return torch.from_numpy(a.astype(np.float32)).to(DEV)


@pytest.mark.parametrize("N,H,W", [(2, 64, 96), (1, 30, 70), (1, 2, 2)])
def test_gather_variants(monkeypatch, N, H, W):
    rng = np.random.RandomState(H * W)
    U, c0 = _t(rng.randn(9, N, H // 2, W // 2, 8)), _t(rng.randn(N, H, W, 8))
    wc, bt = _t(rng.randn(9, 8, 8) / 8), _t(rng.randn(9, 8))

    def run():
        out = torch.full((N, H, W, 8), float("nan"), device=DEV)
        _lib.check(_lib.load().mvster_fpn_out4_gather_f32(capi._ptr(U), 8, capi._ptr(c0), capi._ptr(wc), capi._ptr(bt), capi._ptr(out),
                                                          N, H, W, capi._stream()), "gather")
        torch.cuda.synchronize()
        return out
    monkeypatch.setenv("MVSTER_FPN_GATHER", "1")
    want = run()
    scale = want.abs().max().item()
    assert torch.isfinite(want).all()
    for variant in ("2",):
        monkeypatch.setenv("MVSTER_FPN_GATHER", variant)
        got = run()
        assert torch.isfinite(got).all() and (got - want).abs().max().item() <= 2e-6 * scale, variant


@pytest.mark.parametrize("N,H,W,CL", [(2, 32, 48, 32), (1, 64, 80, 16), (1, 6, 10, 16), (1, 4, 30, 8)])
def test_merge_variants(monkeypatch, N, H, W, CL):
    rng = np.random.RandomState(H * W + CL)
    top, lat = _t(rng.randn(N, H // 2, W // 2, 64)), _t(rng.randn(N, H, W, CL))
    w, bias = _t(rng.randn(CL, 64) / 4), _t(rng.randn(64))
    monkeypatch.setenv("MVSTER_FPN_MERGE", "1")
    want = fpn_engine._merge(top, lat, w, bias)
    torch.cuda.synchronize()
    scale = want.abs().max().item()
    for variant in ("2", "3"):
        monkeypatch.setenv("MVSTER_FPN_MERGE", variant)
        got = fpn_engine._merge(top, lat, w, bias)
        torch.cuda.synchronize()
        assert torch.isfinite(got).all() and (got - want).abs().max().item() <= 2e-6 * scale, variant


@pytest.mark.parametrize("N,H,W", [(5, 64, 128), (1, 5, 12), (2, 3, 4)])
def test_four_pixel_stem(monkeypatch, N, H, W):
    rng = np.random.RandomState(N + H * W)
    img, wt, bias = _t(rng.rand(N, 3, H, W)), _t(rng.randn(9, 3, 8) / 3), _t(rng.randn(8))

    def run():
        out = torch.full((N, H, W, 8), float("nan"), device=DEV)
        _lib.check(_lib.load().mvster_conv_first_f32(capi._ptr(img), capi._ptr(wt), capi._ptr(bias), capi._ptr(out), N, H, W, capi._stream()), "stem")
        torch.cuda.synchronize()
        return out
    monkeypatch.setenv("MVSTER_CONV_FIRST", "1")
    want = run()
    monkeypatch.setenv("MVSTER_CONV_FIRST", "2")
    got = run()
    assert torch.isfinite(want).all() and torch.equal(got, want)          # same FMA chain per output channel


@pytest.mark.parametrize("B,D,H,W,G", [(1, 4, 64, 96, 4), (2, 8, 16, 24, 8), (1, 1, 3, 4, 4)])
def test_four_voxel_conv0(monkeypatch, B, D, H, W, G):
    rng = np.random.RandomState(B * 100 + H * W + G)
    x, wt, bias = _t(rng.randn(B, D, H, W, G)), _t(rng.randn(9, G, 8) / 3), _t(rng.randn(8))
    monkeypatch.setenv("MVSTER_CONV0_PX4", "0")
    want = capi.conv3d_ndhwc(x, wt, bias, 1)
    torch.cuda.synchronize()
    monkeypatch.setenv("MVSTER_CONV0_PX4", "1")
    got = capi.conv3d_ndhwc(x, wt, bias, 1)
    torch.cuda.synchronize()
    assert torch.isfinite(want).all() and torch.equal(got, want)


def test_forward_with_every_variant_on(monkeypatch):
    """Whole forward (default engine) with the switches on against the default kernels: feature-level agreement shows up as
    the same depth map up to the rare near-tie."""
    imgs, proj, dv = synth.make_inputs(1, 3, 128, 192, seed=9)
    imgs, proj, dv = [i.to(DEV) for i in imgs], {k: v.to(DEV) for k, v in proj.items()}, dv.to(DEV)
    model = build_model(SHIPPED, seed=5).to(DEV).eval()
    model.use_cuda_graph = False  # the switches are read at launch time: a replayed graph would keep the default kernels
    with torch.no_grad():
        want = {k: v.clone() for k, v in model(imgs, proj, dv)["stage4"].items() if isinstance(v, torch.Tensor)}
        for k, v in (("MVSTER_FPN_GATHER", "1"), ("MVSTER_FPN_MERGE", "1"), ("MVSTER_CONV_FIRST", "1"), ("MVSTER_CONV0_PX4", "0")):  # the first-generation kernels
            monkeypatch.setenv(k, v)
        n0 = _lib.launch_count()
        got = model(imgs, proj, dv)["stage4"]
        assert _lib.launch_count() > n0
    torch.cuda.synchronize()
    same = ((got["depth"] - want["depth"]).abs() <= 1e-4 * want["depth"].abs()).float().mean().item()
    assert same > 0.97, same                                       # CPU emulation, 64x128 golden frame: every pixel
    assert (got["attn_weight"] - want["attn_weight"]).abs().median().item() < 1e-4
