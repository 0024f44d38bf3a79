"""Random-shape fuzzing of the SIMT kernels on the CPU emulation (manual tool, not collected by pytest).  Meant to be run
with the AddressSanitizer build so that any out-of-bounds access at an odd shape is reported:

    MVSTER_EMU_ASAN=1 LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:verify_asan_link_order=0 \\
        PYTHONPATH=.:tests python tests/emu/fuzz.py [iterations]

Part 1: every default-off kernel variant (gather 2 / 3, merge 2 / 3, four-pixel stem, four-voxel conv0) against the default
kernel at random sizes (ragged and tiny ones included).  Part 2: the warp + ET forward (library's choice of kernel) and the
backward kernel against the oracle / autograd through the oracle at random channel configurations, view counts, baselines
and hypothesis ranges.  Part 3 (round 2): the bf16-storage kernels - cast, bf16 warp + ET (tiled / window, natural and
interleaved channels) against the oracle on bf16-valued features - at random shapes (the regulariser's conv0 kernel with its
bf16 / packed fp16-pair stores has no entry point of its own: tests/test_emu_bf16.py drives it through the reg2d entry points)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path[:0] = [str(HERE), str(HERE.parent), str(HERE.parents[1])]
import install  # noqa: E402

install.install()
from util import oracle  # noqa: E402
from mvster_b200 import _lib, capi, fpn_engine, synth, train_ops  # noqa: E402


def t(a):
    return torch.from_numpy(a.astype(np.float32))


def setenv(k, v):
    if v is None:
        os.environ.pop(k, None)
    else:
        os.environ[k] = v


def variants(iters, rng):
    lib = _lib.load()
    worst = {"merge3": 0.0}
    for _ in range(iters):
        N, H, W = int(rng.randint(1, 3)), 2 * int(rng.randint(1, 20)), 2 * int(rng.randint(1, 45))
        U, c0 = t(rng.randn(9, N, H // 2, W // 2, 8)), t(rng.randn(N, H, W, 8))
        wc, bt = t(rng.randn(9, 8, 8) / 8), t(rng.randn(9, 8))
        outs = {}
        for v in ("1", "2"):
            setenv("MVSTER_FPN_GATHER", v)
            out = torch.full((N, H, W, 8), float("nan"))
            assert lib.mvster_fpn_out4_gather_f32(capi._ptr(U), 8, capi._ptr(c0), capi._ptr(wc), capi._ptr(bt), capi._ptr(out), N, H, W, None) == 0
            outs[v] = out
        setenv("MVSTER_FPN_GATHER", None)
        assert torch.isfinite(outs["1"]).all() and torch.equal(outs["2"], outs["1"]), ("gather", N, H, W)
        CL = int(rng.choice([8, 16, 32]))
        top, lat = t(rng.randn(N, H // 2, W // 2, 64)), t(rng.randn(N, H, W, CL))
        w, b = t(rng.randn(CL, 64) / 4), t(rng.randn(64))
        outs = {}
        for v in ("1", "2", "3"):
            setenv("MVSTER_FPN_MERGE", v)
            outs[v] = fpn_engine._merge(top, lat, w, b)
        setenv("MVSTER_FPN_MERGE", None)
        e = ((outs["3"] - outs["1"]).abs().max() / outs["1"].abs().max()).item()
        assert torch.equal(outs["2"], outs["1"]) and torch.isfinite(outs["3"]).all() and e < 2e-6, ("merge", N, H, W, CL, e)
        worst["merge3"] = max(worst["merge3"], e)
        Hs, Ws = int(rng.randint(1, 12)), int(rng.randint(1, 40))
        img, wt, bs = t(rng.rand(N, 3, Hs, Ws)), t(rng.randn(9, 3, 8) / 3), t(rng.randn(8))
        outs = {}
        for v in (None, "2"):
            setenv("MVSTER_CONV_FIRST", v)
            out = torch.full((N, Hs, Ws, 8), float("nan"))
            assert lib.mvster_conv_first_f32(capi._ptr(img), capi._ptr(wt), capi._ptr(bs), capi._ptr(out), N, Hs, Ws, None) == 0
            outs[v] = out
        setenv("MVSTER_CONV_FIRST", None)
        assert torch.equal(outs["2"], outs[None]), ("stem", N, Hs, Ws)
        G, D, Wv = int(rng.choice([4, 8])), int(rng.randint(1, 4)), 2 * int(rng.randint(1, 20))
        x, w0, b0 = t(rng.randn(N, D, Hs, Wv, G)), t(rng.randn(9, G, 8) / 3), t(rng.randn(8))
        outs = {}
        for v in ("0", "1"):
            setenv("MVSTER_CONV0_PX4", v)
            outs[v] = capi.conv3d_ndhwc(x, w0, b0, 1)
        setenv("MVSTER_CONV0_PX4", None)
        assert torch.equal(outs["1"], outs["0"]), ("conv0", N, D, Hs, Wv, G)
    return worst


def warp_et(iters, rng):
    worst_f = worst_b = 0.0
    cfgs = [(64, 8, 8), (32, 8, 8), (16, 4, 4), (8, 4, 4), (16, 8, 4), (8, 8, 8), (32, 4, 8)]
    for it in range(iters):
        C_, G, D = cfgs[it % len(cfgs)]
        B, nv, H, W = int(rng.randint(1, 3)), int(rng.randint(2, 5)), int(rng.randint(2, 9)), int(rng.randint(2, 33))
        step = float(rng.choice([0.3, 1.0, 5.0, 20.0]))
        feats = [t(rng.randn(B, C_, H, W)) for _ in range(nv)]
        cams = synth.stage_projections(synth.arc_cameras(nv, H, W, step), B, num_stage=1)["stage1"]
        base = t(np.exp(rng.uniform(np.log(50.0), np.log(5e3), (B, 1, H, W))))
        hypo = (base * torch.linspace(1.0, float(rng.choice([0.999, 0.9, 0.3])), D).reshape(1, D, 1, 1)).contiguous()
        gout = t(rng.randn(B, G, D, H, W))
        ref = [f.clone().requires_grad_(True) for f in feats]
        want_c = oracle.et_aggregate(ref, cams, hypo, True, G, 2.0)
        want_g = torch.autograd.grad(want_c, ref, gout)
        lv = [f.clone().requires_grad_(True) for f in feats]
        c = train_ops.aggregate(lv, cams, hypo, G, 2.0)
        g = torch.autograd.grad(c, lv, gout)
        ef = (c.detach() - want_c.detach()).abs().max().item() / max(want_c.abs().max().item(), 1e-6)
        eb = max((g[v] - want_g[v]).abs().max().item() / max(want_g[v].abs().max().item(), 1e-6) for v in range(nv))
        assert torch.isfinite(c).all() and ef < 1e-3 and eb < 2e-3, (C_, G, D, B, nv, H, W, step, ef, eb)
        worst_f, worst_b = max(worst_f, ef), max(worst_b, eb)
    return worst_f, worst_b


def bf16_storage(iters, rng):
    q = lambda x: x.to(torch.bfloat16).float()
    worst_eq = 1.0
    cfgs = [(64, 8, 8), (32, 8, 8), (16, 4, 4), (8, 4, 4)]
    for it in range(iters):
        n = int(rng.randint(1, 300))
        x = t(rng.randn(n) * 10 ** rng.uniform(-3, 3, n))
        assert torch.equal(capi.cast_bf16(x).view(torch.int16), x.to(torch.bfloat16).view(torch.int16)), ("cast", n)
        C_, G, D = cfgs[it % len(cfgs)]
        B, nv, H, W = int(rng.randint(1, 3)), int(rng.randint(2, 5)), int(rng.randint(2, 9)), int(rng.randint(2, 40))
        step = float(rng.choice([0.3, 1.0, 5.0]))
        feats = [q(t(rng.randn(B, C_, H, W))) for _ in range(nv)]
        cams = synth.stage_projections(synth.arc_cameras(nv, H, W, step), B, num_stage=1)["stage1"]
        base = t(np.exp(rng.uniform(np.log(300.0), np.log(2e3), (B, 1, H, W))))
        hypo = (base * torch.linspace(1.0, float(rng.choice([0.999, 0.97, 0.5])), D).reshape(1, D, 1, 1)).contiguous()
        want = q(oracle.et_aggregate(feats, cams, hypo, True, G, 2.0))
        nh = [f.permute(0, 2, 3, 1).contiguous() for f in feats]
        for window, il in ((None, False), (False, False), (True, C_ // G in (2, 4))):
            perm = capi.interleave_perm(C_, G) if il else list(range(C_))
            fb = [f[..., perm].contiguous().to(torch.bfloat16) for f in nh]
            got = capi.et_fuse_bf16(fb[0], fb[1:], capi.pose(cams), hypo, G, 2.0, window=window, interleaved=il).float().permute(0, 4, 1, 2, 3)
            scale = max(want.abs().max().item(), 1e-6)
            assert torch.isfinite(got).all() and (got - want).abs().max().item() <= (2.0 ** -7 + 1e-3) * scale, (C_, G, D, B, nv, H, W, step, window, il)
            worst_eq = min(worst_eq, (got == want).float().mean().item())
    return worst_eq


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.RandomState(2024)
    print("variants vs default kernels, worst deviation / max:", variants(n, rng))
    print("warp + ET forward / backward vs oracle, worst deviation / max:", warp_et(n, rng))
    print("bf16 storage: cast bit-exact, warp + ET within one bf16 spacing of the rounded oracle; smallest fraction of exactly equal values:",
          bf16_storage(n, rng))
