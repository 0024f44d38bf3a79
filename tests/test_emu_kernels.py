"""The SIMT half of libmvster_b200 executed on the CPU and checked against the oracle - kernel logic coverage without a GPU.

tests/emu/build_emu.py compiles the SOURCE files of mvster_b200/csrc that contain no tensor-core code (warp + ET kernels in
all variants, CUDA-core convolutions / reg2d / reg3d, head, hypotheses, pose, FPN glue, geometric-consistency filter) with
g++ against tests/emu/simt_emu.h: one OS thread per CUDA thread of a block, real barriers for __syncthreads / warp votes /
shuffles, `<<<...>>>` launches rewritten to a host loop over blocks.  The result exports the same C ABI and takes host
pointers; here `mvster_b200.capi` is pointed at it (test-only monkeypatching: the product has no such path and still raises
without a GPU), so the very wrappers and kernels the GPU tests exercise run on small inputs.  The GPU object code is
byte-identical with and without the few `#ifdef MVSTER_CPU_EMU` twins of inline-PTX helpers (SASS hashes compared)."""
import ctypes as C
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from util import GOLDEN, REPO, SHIPPED, build_model, narrow_et_inputs, oracle
from oracle import fusion_oracle

from mvster_b200 import _lib, capi, fpn_engine, packing, synth

sys.path.insert(0, str(REPO / "tests" / "emu"))


@pytest.fixture(scope="module")
def emu_lib():
    import install
    return install.load()


@pytest.fixture()
def emu(emu_lib, monkeypatch):
    """capi / fpn_engine on the emulation library, CPU tensors accepted."""
    import install
    monkeypatch.setattr(_lib, "_lib", emu_lib)
    monkeypatch.setattr(capi, "_chk", install.cpu_chk)
    monkeypatch.setattr(capi, "_stream", lambda: C.c_void_p(0))
    # every buffer the wrappers allocate starts as NaN: an output element a kernel does not write cannot pass a comparison
    real_empty = torch.empty

    def nan_empty(*a, **kw):
        t = real_empty(*a, **kw)
        return t.fill_(float("nan")) if t.is_floating_point() else t
    monkeypatch.setattr(torch, "empty", nan_empty)
    return emu_lib


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def from_ndhwc(t):
    return t.permute(0, 4, 1, 2, 3).contiguous()


# ----------------------------------------------------------------------------- hypotheses, pose
def test_hypotheses_and_pose_on_cpu(emu):
    dv = torch.tensor([[425.0, 700.0, 935.0], [300.0, 500.0, 800.0]])
    got = capi.hypo_init_inverse(dv, 8, 8, 12)
    assert torch.equal(got, oracle.hypo_init_inverse(dv, 8, 8, 12))                       # bit-exact, as on the GPU
    rng = np.random.RandomState(0)
    inv_max = torch.from_numpy(rng.uniform(1 / 900, 1 / 600, (2, 8, 12)).astype(np.float32))
    inv_min = inv_max + torch.from_numpy(rng.uniform(1e-5, 3e-4, (2, 8, 12)).astype(np.float32))
    for D in (4, 8):
        got = capi.hypo_schedule_inverse(inv_min, inv_max, D, 16, 24)
        want = oracle.hypo_schedule_inverse(inv_min, inv_max, D, 16, 24)
        assert ((got - want).abs() / want.abs()).max().item() < 5e-7
    _, proj, _ = synth.make_inputs(2, 4, 64, 64, seed=1, step_deg=5.0)
    cams = proj["stage3"]
    got = capi.pose(cams)
    ref_full = oracle.compose_projection(cams[:, 0])
    for v in range(1, 4):
        R, t = oracle.relative_pose(oracle.compose_projection(cams[:, v]), ref_full)
        want = torch.cat([R.reshape(2, 9), t.reshape(2, 3)], 1)
        assert ((got[:, v - 1] - want).abs().max() / want.abs().max()).item() < 2e-6


# ----------------------------------------------------------------------------- warp + ET kernels, every variant
ET_SMALL = [  # (B, nv, C, G, D, H, W, step_deg): wide hypothesis ranges -> per-hypothesis kernels; the window kernel falls back
    (1, 3, 64, 8, 8, 4, 8, 1.0),
    (1, 3, 32, 8, 8, 8, 8, 5.0),
    (1, 2, 16, 4, 4, 8, 16, 2.0),
    (1, 3, 8, 4, 4, 8, 32, 1.0),
    (1, 2, 16, 8, 4, 4, 8, 3.0),   # generic kernel only: C/G = 2, G = 8
]


def et_inputs(B, nv, C, G, D, H, W, step, seed=0):
    rng = np.random.RandomState(seed)
    feats = [torch.from_numpy(rng.randn(B, C, H, W).astype(np.float32)) for _ in range(nv)]
    cams = synth.stage_projections(synth.arc_cameras(nv, H, W, step), B, num_stage=1)["stage1"]
    hypo = oracle.hypo_init_inverse(torch.tensor([[425.0, 935.0]] * B), D, H, W)
    hypo = hypo * torch.from_numpy(rng.uniform(0.97, 1.03, (B, D, H, W)).astype(np.float32))
    return feats, cams, hypo


@pytest.mark.parametrize("case", ET_SMALL)
def test_et_kernels_on_cpu_match_oracle(emu, case):
    B, nv, C_, G, D, H, W, step = case
    feats, cams, hypo = et_inputs(*case)
    want = oracle.et_aggregate(feats, cams, hypo, True, G, 2.0)
    scale = want.abs().max().item()
    ref, srcs, pose = nhwc(feats[0]), [nhwc(f) for f in feats[1:]], capi.pose(cams)
    variants = [dict(generic=True), dict(window=False), dict(window=True)]
    for kw in variants:
        got = from_ndhwc(capi.et_fuse(ref, srcs, pose, hypo, G, 2.0, **kw))
        assert (got - want).abs().max().item() <= 2e-4 * scale, kw


WIN_SMALL = [  # (B, nv, C, G, D, H, W, step_deg, rel_span): hypotheses within one or two source cells -> window path
    (1, 3, 8, 4, 4, 8, 64, 1.0, 0.3),
    (1, 3, 8, 4, 4, 6, 40, 3.0, 0.12),
    (1, 3, 16, 4, 4, 8, 32, 2.0, 0.2),
    (2, 2, 32, 8, 8, 4, 16, 4.0, 0.06),
]


@pytest.mark.parametrize("case", WIN_SMALL)
def test_window_kernel_on_cpu_matches_oracle(emu, case):
    B, nv, C_, G, D, H, W, step, span = case
    feats, cams, hypo = narrow_et_inputs(B, nv, C_, D, H, W, step, span, seed=13)
    want = oracle.et_aggregate(feats, cams, hypo, True, G, 2.0)
    for tma in (False, True):  # taps gathered from global memory / from TMA-staged boxes in shared memory
        got = from_ndhwc(capi.et_fuse(nhwc(feats[0]), [nhwc(f) for f in feats[1:]], capi.pose(cams), hypo, G, 2.0, window=True, tma=tma))
        assert ("tma" in capi.et_last_kernel()) == tma, capi.et_last_kernel()
        assert torch.isfinite(got).all() and (got - want).abs().max().item() <= 2e-4 * want.abs().max().item(), tma


TMA_CASES = [  # (B, nv, C, G, D, H, W, step_deg, rel_span): several tiles per CTA, ragged tiles, batch > 1, more views than box slots
    (1, 6, 8, 4, 4, 37, 70, 1.0, 0.12),
    (2, 3, 8, 4, 4, 16, 40, 2.0, 0.3),
    (1, 5, 16, 4, 4, 17, 48, 1.5, 0.2),
    (2, 4, 32, 8, 8, 7, 33, 3.0, 0.06),
]


@pytest.mark.parametrize("case", TMA_CASES)
def test_tma_staged_kernel_on_cpu_matches_oracle(emu, case, monkeypatch):
    """et_fuse_tma_kernel with the mbarrier / TMA primitives replaced by host twins (same protocol: producer warp, slot ring,
    swizzled boxes with zero fill): against the oracle, against the global-gather window kernel, in partial + accumulate mode,
    and with 7-row tiles (two CTAs per SM on the GPU)."""
    B, nv, C_, G, D, H, W, step, span = case
    feats, cams, hypo = narrow_et_inputs(B, nv, C_, D, H, W, step, span, seed=17)
    # a few pixels with hypotheses far outside the narrow band: those warps must take the global-gather path
    hypo[:, :, 0, :3] *= torch.tensor([1.0, 0.8, 0.6, 0.4] * (D // 4)).reshape(1, D, 1)
    want = oracle.et_aggregate(feats, cams, hypo, True, G, 2.0)
    scale = want.abs().max().item()
    ref, srcs, pose = nhwc(feats[0]), [nhwc(f) for f in feats[1:]], capi.pose(cams)
    got = from_ndhwc(capi.et_fuse(ref, srcs, pose, hypo, G, 2.0, tma=True))
    assert "tma" in capi.et_last_kernel()
    assert torch.isfinite(got).all() and (got - want).abs().max().item() <= 2e-4 * scale
    win = from_ndhwc(capi.et_fuse(ref, srcs, pose, hypo, G, 2.0, tma=False))
    assert (got - win).abs().max().item() <= 2e-5 * scale
    # group-interleaved channel layout (MVSTER_ET_INTERLEAVED): same features with permuted channels, both kernels
    perm = capi.interleave_perm(C_, G)
    ref_il, srcs_il = ref[..., perm].contiguous(), [s_[..., perm].contiguous() for s_ in srcs]
    for tma in (True, False):
        got_il = from_ndhwc(capi.et_fuse(ref_il, srcs_il, pose, hypo, G, 2.0, tma=tma, interleaved=True))
        assert "interleaved" in capi.et_last_kernel() or tma
        assert (got_il - got).abs().max().item() <= 5e-6 * scale, tma
    # view sharding: partial sums of the first views, then accumulate the rest, then normalise
    k = max(1, (nv - 1) // 2)
    wsum = torch.empty(B, D, H, W)
    cost = capi.et_fuse(ref, srcs[:k], pose[:, :k].contiguous(), hypo, G, 2.0, wsum=wsum, partial=True, tma=True)
    if k < nv - 1:
        capi.et_fuse(ref, srcs[k:], pose[:, k:].contiguous(), hypo, G, 2.0, cost=cost, wsum=wsum, partial=True, accumulate=True, tma=True)
    capi.et_normalize(cost, wsum)
    assert (from_ndhwc(cost) - want).abs().max().item() <= 2e-4 * scale
    if C_ == 8:
        monkeypatch.setenv("MVSTER_ET_TMA_TH", "15")
        got15 = from_ndhwc(capi.et_fuse(ref, srcs, pose, hypo, G, 2.0, tma=True))
        assert "15x32" in capi.et_last_kernel()
        assert (got15 - want).abs().max().item() <= 2e-4 * scale


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_et_kernels_agree_on_extreme_geometry_on_cpu(emu, seed):
    """Wide baselines and hypotheses from very near to very far: most taps fall outside the source images, sampling positions
    reach thousands of pixels, hypotheses of one pixel straddle several cells - the window kernel must fall back where it has to
    and every variant must still match the oracle (zeros padding per tap)."""
    rng = np.random.RandomState(100 + seed)
    B, nv, C_, G, D, H, W = 1, 3, 8, 4, 4, 8, 32
    feats = [torch.from_numpy(rng.randn(B, C_, H, W).astype(np.float32)) for _ in range(nv)]
    cams = synth.stage_projections(synth.arc_cameras(nv, H, W, [25.0, 8.0, 1.0][seed]), B, num_stage=1)["stage1"]
    base = torch.from_numpy(np.exp(rng.uniform(np.log(5.0), np.log(5e4), (B, 1, H, W))).astype(np.float32))
    hypo = (base * torch.tensor([1.0, 0.999, 0.5, 0.01]).reshape(1, D, 1, 1)).contiguous()
    want = oracle.et_aggregate(feats, cams, hypo, True, G, 2.0)
    ref, srcs, pose = nhwc(feats[0]), [nhwc(f) for f in feats[1:]], capi.pose(cams)
    scale = max(want.abs().max().item(), 1e-3)
    for kw in (dict(generic=True), dict(window=False), dict(window=True)):
        got = from_ndhwc(capi.et_fuse(ref, srcs, pose, hypo, G, 2.0, **kw))
        assert torch.isfinite(got).all(), kw
        assert (got - want).abs().max().item() <= 5e-4 * scale, (kw, (got - want).abs().max().item(), scale)


ET_BWD = [  # (B, nv, C, G, D, H, W, step_deg): every (C/G, G, D) shape of the shipped stages plus C/G = 1 and 2 with G = 8
    (1, 3, 64, 8, 8, 4, 8, 1.0),
    (2, 3, 32, 8, 8, 4, 8, 5.0),
    (1, 2, 16, 4, 4, 8, 16, 2.0),
    (1, 4, 8, 4, 4, 8, 16, 1.0),
    (1, 2, 16, 8, 4, 4, 8, 3.0),
    (1, 2, 8, 8, 8, 4, 8, 3.0),
]


@pytest.mark.parametrize("case", ET_BWD)
def test_et_backward_kernel_on_cpu_matches_autograd(emu, case):
    """mvster_et_fuse_bwd_f32 through the autograd node of train_ops against torch autograd through the oracle's PyTorch
    formulation (fp64 ops on the same inputs): gradients w.r.t. the reference and every source feature map."""
    from mvster_b200 import train_ops
    B, nv, C_, G, D, H, W, step = case
    feats, cams, hypo = et_inputs(*case, seed=5)
    rng = np.random.RandomState(7)
    gout = torch.from_numpy(rng.randn(B, G, D, H, W).astype(np.float32))
    f64 = [f.double().requires_grad_(True) for f in feats]
    want_cost = oracle.et_aggregate(f64, cams.double(), hypo.double(), True, G, 2.0)
    want = torch.autograd.grad(want_cost, f64, gout.double())
    f32 = [f.clone().requires_grad_(True) for f in feats]
    cost = train_ops.aggregate(f32, cams, hypo, G, 2.0)
    assert cost.shape == (B, G, D, H, W)
    assert (cost.detach() - want_cost.detach().float()).abs().max().item() <= 2e-4 * want_cost.abs().max().item()
    got = torch.autograd.grad(cost, f32, gout)
    for v in range(nv):
        scale = want[v].abs().max().item()
        assert scale > 0
        assert (got[v] - want[v].float()).abs().max().item() <= 2e-4 * scale, (v, (got[v] - want[v].float()).abs().max().item(), scale)


@pytest.mark.parametrize("name", ["c8", "c16", "c32", "c64"])
def test_et_backward_kernel_on_cpu_matches_reference_gradients(emu, name):
    """... and against gradients taken through the unmodified reference (tests/golden/et_backward.npz)."""
    from mvster_b200 import train_ops
    from test_et_backward_golden import load_case
    c = load_case(name)
    leaves = [f.clone().requires_grad_(True) for f in c["feats"]]
    cost = train_ops.aggregate(leaves, c["cams"], c["hypo"], c["G"], 2.0)
    assert (cost.detach() - c["cost"]).abs().max().item() <= 2e-4 * c["cost"].abs().max().item()
    grads = torch.autograd.grad(cost, leaves, c["gout"])
    for v in range(c["nv"]):
        scale = c["grads"][v].abs().max().item()
        err = (grads[v] - c["grads"][v]).abs().max().item()
        assert err <= 2e-4 * scale, (v, err, scale)


def test_et_backward_skips_views_without_grad_on_cpu(emu):
    from mvster_b200 import train_ops
    feats, cams, hypo = et_inputs(1, 3, 8, 4, 4, 8, 16, 2.0, seed=9)
    gout = torch.ones(1, 4, 4, 8, 16)
    full = [f.clone().requires_grad_(True) for f in feats]
    g_full = torch.autograd.grad(train_ops.aggregate(full, cams, hypo, 4, 2.0), full, gout)
    some = [f.clone().requires_grad_(i != 1) for i, f in enumerate(feats)]           # source view 1 frozen
    g_some = torch.autograd.grad(train_ops.aggregate(some, cams, hypo, 4, 2.0), [some[0], some[2]], gout)
    assert torch.allclose(g_some[0], g_full[0], rtol=0, atol=1e-6 * g_full[0].abs().max().item())
    assert torch.allclose(g_some[1], g_full[2], rtol=0, atol=1e-6 * g_full[2].abs().max().item())


def test_training_step_with_kernels_on_cpu_matches_pytorch_ops(emu, monkeypatch):
    """MVS4net in train mode (mono decoder, OT + L1 loss): parameter gradients with the aggregation on the fused kernels
    (forward + backward, emulated) against the all-PyTorch formulation."""
    from mvster_b200 import MVS4net_loss, train_ops
    monkeypatch.setattr(train_ops, "usable", lambda feats, G=0, D=0: True)
    model = build_model(SHIPPED, seed=3).train()
    for m in model.modules():
        if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
            m.eval()
    imgs, proj, dv = synth.make_inputs(1, 3, 64, 64, seed=4, step_deg=1.0)
    gt = {f"stage{k + 1}": torch.full((1, 64 >> (3 - k), 64 >> (3 - k)), 680.0) for k in range(4)}
    mask = {k: torch.ones_like(v) for k, v in gt.items()}
    grads, losses = {}, {}
    for use in (False, True):
        model.stagenet.train_et = use
        model.zero_grad(set_to_none=True)
        n0 = _lib.launch_count()
        out = model(imgs, proj, dv)
        loss = MVS4net_loss(out, gt, mask, stage_lw=[1, 1, 1, 1], l1ot_lw=[1, 1], inverse_depth=True, mono=True)[0]
        loss.backward()
        n = _lib.launch_count() - n0                                # per stage: pose, [tile boxes,] forward, normalise, backward
        assert (16 <= n <= 20) if use else n == 0
        losses[use] = loss.item()
        grads[use] = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    assert abs(losses[True] - losses[False]) <= 1e-5 * abs(losses[False])
    assert grads[True].keys() == grads[False].keys() and len(grads[True]) > 50
    # the two forwards differ by fp32 rounding (attn_weight up to 2e-3 at stage 4 with these random weights, identical argmax),
    # which bounds how closely the gradients can agree; prob.bias has a zero true gradient (softmax shift invariance) - skipped
    top = max(g.abs().max().item() for g in grads[False].values())
    worst = max(((grads[True][n] - grads[False][n]).abs().max() / grads[False][n].abs().max()).item()
                for n in grads[True] if grads[False][n].abs().max().item() > 1e-5 * top)
    assert worst < 5e-2, worst                                      # measured on the CPU emulation: 7.6e-3


def test_warp_edge_cases_of_the_reference_through_the_kernels_on_cpu(emu):
    from util import warp_edge_through_et
    outs, want = warp_edge_through_et(capi)
    assert (want == 0).float().mean() > 0.2
    for got in outs:
        assert torch.isfinite(got).all()
        assert (got - want).abs().max().item() <= 2e-6 * want.abs().max().item()
        assert ((got == 0) == (want == 0)).float().mean() > 0.999             # the zero padding is reproduced tap for tap


def test_more_source_views_than_one_launch_takes_on_cpu(emu):
    """19 source views > MVSTER_MAX_VIEWS = 16: the forward chains two launches through the partial sums, the backward adds the
    reference-feature gradients of the two chunks; both against the oracle (fp64 autograd for the gradients)."""
    from mvster_b200 import train_ops
    assert capi.MAX_VIEWS == 16
    feats, cams, hypo = et_inputs(1, 20, 8, 4, 4, 4, 8, 0.3, seed=17)
    want = oracle.et_aggregate(feats, cams, hypo, True, 4, 2.0)
    n0 = _lib.launch_count()
    got = from_ndhwc(capi.et_fuse(nhwc(feats[0]), [nhwc(f) for f in feats[1:]], capi.pose(cams), hypo, 4, 2.0))
    assert 4 <= _lib.launch_count() - n0 <= 6                             # pose, two chained launches (+ their tile-box launches), normalise
    assert (got - want).abs().max().item() <= 2e-4 * want.abs().max().item()
    gout = torch.from_numpy(np.random.RandomState(1).randn(1, 4, 4, 4, 8).astype(np.float32))
    f64 = [f.double().requires_grad_(True) for f in feats]
    gwant = torch.autograd.grad(oracle.et_aggregate(f64, cams.double(), hypo.double(), True, 4, 2.0), f64, gout.double())
    leaves = [f.clone().requires_grad_(True) for f in feats]
    ggot = torch.autograd.grad(train_ops.aggregate(leaves, cams, hypo, 4, 2.0), leaves, gout)
    for v in range(20):
        assert (ggot[v] - gwant[v].float()).abs().max().item() <= 2e-4 * gwant[v].abs().max().item(), v


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_et_backward_on_extreme_geometry_on_cpu(emu, seed):
    """The inputs of test_et_kernels_agree_on_extreme_geometry_on_cpu (most taps outside the source images, positions in the
    thousands of pixels, near and far hypotheses): the backward kernel scatters only into valid taps and matches fp32 autograd
    through the oracle (same operation order for the sampling positions, hence the same taps)."""
    from mvster_b200 import train_ops
    rng = np.random.RandomState(100 + seed)
    B, nv, C_, G, D, H, W = 1, 3, 8, 4, 4, 8, 32
    feats = [torch.from_numpy(rng.randn(B, C_, H, W).astype(np.float32)) for _ in range(nv)]
    cams = synth.stage_projections(synth.arc_cameras(nv, H, W, [25.0, 8.0, 1.0][seed]), B, num_stage=1)["stage1"]
    base = torch.from_numpy(np.exp(rng.uniform(np.log(5.0), np.log(5e4), (B, 1, H, W))).astype(np.float32))
    hypo = (base * torch.tensor([1.0, 0.999, 0.5, 0.01]).reshape(1, D, 1, 1)).contiguous()
    gout = torch.from_numpy(rng.randn(B, G, D, H, W).astype(np.float32))
    ref_leaves = [f.clone().requires_grad_(True) for f in feats]
    want = torch.autograd.grad(oracle.et_aggregate(ref_leaves, cams, hypo, True, G, 2.0), ref_leaves, gout)
    leaves = [f.clone().requires_grad_(True) for f in feats]
    got = torch.autograd.grad(train_ops.aggregate(leaves, cams, hypo, G, 2.0), leaves, gout)
    for v in range(nv):
        assert torch.isfinite(got[v]).all()
        scale = max(want[v].abs().max().item(), 1e-6)
        assert (got[v] - want[v]).abs().max().item() <= 1e-3 * scale, (v, (got[v] - want[v]).abs().max().item(), scale)
        assert ((got[v] == 0) == (want[v] == 0)).float().mean().item() > 0.98     # untouched source pixels stay exactly zero


def test_et_variants_sqdiff_no_fuse_d_partial_on_cpu(emu):
    feats, cams, hypo = et_inputs(1, 3, 8, 8, 4, 4, 8, 2.0, seed=2)
    ref, srcs, pose = nhwc(feats[0]), [nhwc(f) for f in feats[1:]], capi.pose(cams)
    want = oracle.et_aggregate(feats, cams, hypo, False, 8, 2.0)                           # group_cor=False: (ref - warped)^2 per channel
    got = from_ndhwc(capi.et_fuse(ref, srcs, pose, hypo, 8, 2.0, group_cor=False))
    assert (got - want).abs().max().item() <= 2e-4 * want.abs().max().item()
    feats, cams, hypo = et_inputs(1, 3, 16, 4, 4, 4, 8, 2.0, seed=3)
    ref, srcs, pose = nhwc(feats[0]), [nhwc(f) for f in feats[1:]], capi.pose(cams)
    want = oracle.et_aggregate(feats, cams, hypo, True, 4, 2.0, attn_fuse_d=False)
    got = from_ndhwc(capi.et_fuse(ref, srcs, pose, hypo, 4, 2.0, fuse_d=False))
    assert (got - want).abs().max().item() <= 2e-4 * want.abs().max().item()
    full = capi.et_fuse(ref, srcs, pose, hypo, 4, 2.0)
    c1, w1 = torch.empty_like(full), torch.empty((1, 4, 4, 8))
    capi.et_fuse(ref, srcs[:1], pose[:, :1].contiguous(), hypo, 4, 2.0, cost=c1, wsum=w1, partial=True)
    capi.et_fuse(ref, srcs[1:], pose[:, 1:].contiguous(), hypo, 4, 2.0, cost=c1, wsum=w1, partial=True, accumulate=True)
    capi.et_normalize(c1, w1)
    assert (c1 - full).abs().max().item() <= 1e-5 * full.abs().max().item()


# ----------------------------------------------------------------------------- regulariser + head (exact-fp32 CUDA-core path)
@pytest.mark.parametrize("k,D,H,W", [(0, 8, 8, 8), (3, 4, 16, 8)])
def test_reg2d_and_head_on_cpu_match_oracle(emu, k, D, H, W):
    sd = build_model(SHIPPED, 5).state_dict()
    G = SHIPPED["group_cor_dim"][k]
    rng = np.random.RandomState(k)
    cost = torch.from_numpy((rng.randn(1, G, D, H, W) * 0.1).astype(np.float32))
    hypo = oracle.hypo_init_inverse(torch.tensor([[425.0, 935.0]]), D, H, W)
    with torch.no_grad():
        logits = oracle.reg2d_logits(sd, f"reg.{k}", cost)
        want = oracle.depth_head(logits, hypo, k, 0.5, True)
    packed = packing.pack_reg2d(sd, f"reg.{k}", capi.reg2d_layer_table(G))
    feat8 = capi.reg2d(packed["blob"], cost.permute(0, 2, 3, 4, 1).contiguous())
    h = capi.head(hypo, 0.5, feat8=feat8, prob_w=packed["prob_w"], prob_b=packed["prob_b"])
    assert (h["attn_weight"] - want["attn_weight"]).abs().max().item() < 2e-5
    gap = want["attn_weight"].topk(2, dim=1).values
    stable = (gap[:, 0] - gap[:, 1]) > 1e-3
    assert torch.equal(h["depth"][stable], want["depth"][stable])                           # winner-take-all: gathered, so bit-exact
    assert ((h["inverse_min_depth"] - want["inverse_min_depth"]).abs()[stable] <= 1e-6 * want["inverse_min_depth"].abs()[stable]).all()


def test_reg3d_on_cpu_matches_oracle(emu):
    cfg = dict(reg_net="reg3d", group_cor=True, group_cor_dim=[8, 8, 4, 4], inverse_depth=True, attn_temp=2)
    sd = build_model(cfg, 6).state_dict()
    rng = np.random.RandomState(4)
    cost = torch.from_numpy((rng.randn(1, 4, 4, 8, 8) * 0.1).astype(np.float32))
    with torch.no_grad():
        want = oracle.reg3d_logits(sd, "reg.3", cost, 2)
    blob = packing.pack_reg3d(sd, "reg.3", capi.reg3d_layer_table(4, 2))
    got = capi.reg3d(blob, cost.permute(0, 2, 3, 4, 1).contiguous(), 2)
    assert (got - want).abs().max().item() <= 1e-5 * want.abs().max().item()


# ----------------------------------------------------------------------------- feature pyramid (CUDA-core path)
@pytest.mark.parametrize("fused_last", [True, False])
def test_native_fpn_on_cpu_matches_oracle(emu, fused_last):
    sd = build_model(SHIPPED, 4).state_dict()
    imgs, _, _ = synth.make_inputs(1, 2, 16, 24, seed=8)
    x = torch.cat(imgs, 0)
    with torch.no_grad():
        want = oracle.fpn4_features(sd, x)
    got = fpn_engine.run_fpn(fpn_engine.pack_fpn(sd), x, 0, fused_last=fused_last)
    for s in range(1, 5):
        g, w = got[f"stage{s}"].permute(0, 3, 1, 2), want[f"stage{s}"]
        assert g.shape == w.shape and ((g - w).abs().max() / w.abs().max()).item() < 2e-5, s


@pytest.mark.parametrize("N,H,W", [(2, 16, 24), (1, 24, 80), (1, 10, 34), (1, 2, 2), (1, 6, 66)])
def test_tiled_gather_variant_returns_the_same_bits_on_cpu(emu, monkeypatch, N, H, W):
    """fpn_out4_gather2_kernel (shared-memory tiles, MVSTER_FPN_GATHER=2) against fpn_out4_gather_kernel: same expression per
    pixel, so the outputs must be identical - multi-tile, ragged and single-pixel-patch shapes."""
    rng = np.random.RandomState(H * W)
    U = torch.from_numpy(rng.randn(9, N, H // 2, W // 2, 8).astype(np.float32))
    c0 = torch.from_numpy(rng.randn(N, H, W, 8).astype(np.float32))
    wc = torch.from_numpy((rng.randn(9, 8, 8) / 8).astype(np.float32))
    bt = torch.from_numpy(rng.randn(9, 8).astype(np.float32))

    def run():
        out = torch.full((N, H, W, 8), float("nan"))
        rc = emu.mvster_fpn_out4_gather_f32(capi._ptr(U), 8, capi._ptr(c0), capi._ptr(wc), capi._ptr(bt), capi._ptr(out), N, H, W, None)
        assert rc == 0
        return out
    monkeypatch.setenv("MVSTER_FPN_GATHER", "1")
    want = run()
    monkeypatch.setenv("MVSTER_FPN_GATHER", "2")
    got = run()
    assert torch.isfinite(want).all() and torch.equal(got, want)
    scale = want.abs().max().item()
    # both against the definition in float64
    up = F.interpolate(U.double().permute(1, 0, 4, 2, 3).reshape(N, 72, H // 2, W // 2), scale_factor=2, mode="bilinear",
                       align_corners=True).reshape(N, 9, 8, H, W)                                   # up2(U_tap) per tap
    lat = torch.einsum("nyxc,tco->ntoyx", c0.double(), wc.double()) + bt.double().reshape(1, 9, 8, 1, 1)
    term = F.pad(up + lat, (1, 1, 1, 1))                                                            # zero padding of the 3x3 conv
    ref = sum(term[:, ky * 3 + kx, :, ky:ky + H, kx:kx + W] for ky in range(3) for kx in range(3)).permute(0, 2, 3, 1)
    for o in (want, got):
        assert (o.double() - ref).abs().max().item() <= 2e-6 * scale


@pytest.mark.parametrize("N,H,W,CL", [(2, 8, 12, 32), (1, 6, 10, 16), (1, 4, 30, 8), (1, 2, 2, 16), (2, 4, 66, 32)])
def test_four_pixel_merge_variant_returns_the_same_bits_on_cpu(emu, monkeypatch, N, H, W, CL):
    """fpn_merge4_kernel (MVSTER_FPN_MERGE=2) against fpn_merge_kernel, ragged rows included (W not a multiple of 4)."""
    rng = np.random.RandomState(H * W + CL)
    top = torch.from_numpy(rng.randn(N, H // 2, W // 2, 64).astype(np.float32))
    lat = torch.from_numpy(rng.randn(N, H, W, CL).astype(np.float32))
    w = torch.from_numpy((rng.randn(CL, 64) / 4).astype(np.float32))
    bias = torch.from_numpy(rng.randn(64).astype(np.float32))
    monkeypatch.setenv("MVSTER_FPN_MERGE", "1")
    want = fpn_engine._merge(top, lat, w, bias)
    monkeypatch.setenv("MVSTER_FPN_MERGE", "2")
    got = fpn_engine._merge(top, lat, w, bias)
    assert torch.isfinite(want).all() and torch.equal(got, want)
    # variant 3 (four lanes per four pixels, packed FMAs, accumulator seeded with up + bias): equal to rounding, and both
    # against the definition in float64
    monkeypatch.setenv("MVSTER_FPN_MERGE", "3")
    got3 = fpn_engine._merge(top, lat, w, bias)
    scale = want.abs().max().item()
    assert torch.isfinite(got3).all() and (got3 - want).abs().max().item() <= 2e-6 * scale
    ref = F.interpolate(top.double().permute(0, 3, 1, 2), scale_factor=2, mode="bilinear", align_corners=True).permute(0, 2, 3, 1) \
        + lat.double() @ w.double() + bias.double()
    for o in (want, got):
        assert (o.double() - ref).abs().max().item() <= 2e-6 * scale


@pytest.mark.parametrize("N,H,W", [(2, 8, 8), (1, 5, 12), (3, 16, 64), (1, 3, 4), (1, 6, 10)])
def test_four_pixel_stem_variant_returns_the_same_bits_on_cpu(emu, monkeypatch, N, H, W):
    """conv_first4_kernel (MVSTER_CONV_FIRST=2: four pixels per thread, packed FMAs) against conv_first_kernel and against
    F.conv2d; W = 10 is not a multiple of 4 and stays on the one-pixel kernel."""
    rng = np.random.RandomState(N * 100 + H * W)
    img = torch.from_numpy(rng.rand(N, 3, H, W).astype(np.float32))
    wt = torch.from_numpy((rng.randn(8, 3, 3, 3) / 3).astype(np.float32))
    bias = torch.from_numpy(rng.randn(8).astype(np.float32))
    packed = wt.permute(2, 3, 1, 0).reshape(9, 3, 8).contiguous()          # [ky*3+kx][cin][cout]

    def run():
        out = torch.empty(N, H, W, 8)
        rc = emu.mvster_conv_first_f32(C.c_void_p(img.data_ptr()), C.c_void_p(packed.data_ptr()), C.c_void_p(bias.data_ptr()),
                                       C.c_void_p(out.data_ptr()), N, H, W, None)
        assert rc == 0
        return out
    monkeypatch.setenv("MVSTER_CONV_FIRST", "1")
    want = run()
    monkeypatch.setenv("MVSTER_CONV_FIRST", "2")
    got = run()
    assert torch.equal(got, want)
    ref = F.relu(F.conv2d(img, wt, bias, 1, 1)).permute(0, 2, 3, 1)
    assert (want - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()


@pytest.mark.parametrize("B,D,H,W,G,relu", [(1, 4, 8, 16, 4, True), (2, 2, 5, 12, 8, True), (1, 1, 3, 4, 4, False), (1, 4, 16, 64, 8, True),
                                            (1, 2, 6, 10, 4, True)])
def test_four_voxel_conv0_variant_returns_the_same_bits_on_cpu(emu, monkeypatch, B, D, H, W, G, relu):
    """conv0_px4_kernel (MVSTER_CONV0_PX4=1; the regulariser's first layer, (1,3,3) G -> 8) against conv_px2_kernel and against
    F.conv3d; W = 10 is not a multiple of 4 and stays on the two-voxel kernel."""
    rng = np.random.RandomState(B * 1000 + H * W + G)
    x = torch.from_numpy(rng.randn(B, D, H, W, G).astype(np.float32))
    wt = torch.from_numpy((rng.randn(8, G, 1, 3, 3) / 3).astype(np.float32))
    bias = torch.from_numpy(rng.randn(8).astype(np.float32))
    packed = wt[:, :, 0].permute(2, 3, 1, 0).reshape(9, G, 8).contiguous()   # [ky*3+kx][cin][cout]
    monkeypatch.setenv("MVSTER_CONV0_PX4", "0")
    want = capi.conv3d_ndhwc(x, packed, bias, 1, relu=relu)
    monkeypatch.setenv("MVSTER_CONV0_PX4", "1")
    got = capi.conv3d_ndhwc(x, packed, bias, 1, relu=relu)
    assert torch.equal(got, want)
    ref = F.conv3d(x.permute(0, 4, 1, 2, 3), wt, bias, 1, (0, 1, 1))
    ref = (F.relu(ref) if relu else ref).permute(0, 2, 3, 4, 1)
    assert (want - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()


# ----------------------------------------------------------------------------- geometric-consistency filter
@pytest.mark.parametrize("name", ["plane_4v_48x64", "plane_3v_40x56_wide"])
def test_geo_consistency_kernel_on_cpu_matches_reference(emu, name):
    from mvster_b200.fusion import _prepare_mats
    z = np.load(GOLDEN / "fusion" / f"{name}.npz")
    nv, H, W, seed = (int(x) for x in z["meta"])
    views = fusion_oracle.synthetic_scene(nv, H, W, seed, float(z["step"]))
    ref = views[0]
    count, dsum = np.zeros((H, W), np.int32), np.zeros((H, W), np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    for v in range(1, nv):
        s = views[v]
        mats = (C.c_double * 60)(*_prepare_mats(ref["K"], ref["E"], s["K"], s["E"]).tolist())
        dr, xs, ys = (np.empty((H, W), np.float32) for _ in range(3))
        mask = np.empty((H, W), np.uint8)
        rc = emu.mvster_geo_consistency_f32(p(ref["depth"]), p(s["depth"]), mats, p(dr), p(xs), p(ys), p(mask), p(count), p(dsum),
                                            H, W, H, W, 1.0, 0.01, None)
        assert rc == 0
        assert np.array_equal(mask.astype(bool), z[f"mask{v}"]) and np.array_equal(dr, z[f"depth_reprojected{v}"])
        assert np.array_equal(xs, z[f"x2d_src{v}"], equal_nan=True) and np.array_equal(ys, z[f"y2d_src{v}"], equal_nan=True)
    assert np.array_equal(count, z["geo_mask_sum"])
    assert np.array_equal((dsum + ref["depth"]) / (count + 1), z["depth_est_averaged"])


# ----------------------------------------------------------------------------- the whole forward
@pytest.mark.parametrize("name", ["shipped_b1_v3_64x128", "reg3d_b1_v2_64x64", "plain_b1_v2_64x64", "shipped_b1_v3_64x128+variants", "shipped_b2_v2_64x64"])
def test_engine_forward_on_cpu_matches_the_reference_golden(emu, monkeypatch, name):
    """InferenceEngine.forward - native feature pyramid, four cascade stages, every convolution on the exact-fp32 CUDA-core
    kernels - executed on the emulation library with CPU tensors, against the outputs of the unmodified reference
    (tests/golden/shipped_b1_v3_64x128.npz): the same tie-aware, drift-aware criterion as the GPU test, and here every pixel
    of every stage agrees."""
    import contextlib
    from util import GOLDEN_CASES, load_golden, top2_gap
    from mvster_b200.engine import InferenceEngine
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    if name.endswith("+variants"):   # the prepared (opt-in) CUDA-core kernel variants switched on
        name = name[:-len("+variants")]
        for k, v in (("MVSTER_FPN_GATHER", "1"), ("MVSTER_FPN_MERGE", "1"), ("MVSTER_CONV_FIRST", "1"), ("MVSTER_CONV0_PX4", "0")):  # the first-generation kernels
            monkeypatch.setenv(k, v)
    z, imgs, proj, dv = load_golden(name)
    m = build_model(GOLDEN_CASES[name], int(z["meta_seed"]))
    m.reg_precision = m.fpn_precision = "fp32"
    m.overlap_stages, m.fpn_backend = False, "native"
    eng = InferenceEngine(torch.device("cpu"))
    eng.refresh_weights(m)
    with torch.no_grad():
        out = eng.forward(m, imgs, proj, dv)
    drift_free = torch.ones_like(torch.from_numpy(z["s1_depth"]), dtype=torch.bool)
    for s_ in range(1, 5):
        st = out[f"stage{s_}"]
        for key in ("depth", "photometric_confidence", "hypo_depth", "attn_weight", "inverse_min_depth", "inverse_max_depth"):
            if f"s{s_}_{key}" in z.files:  # the inverse-range outputs exist only with inverse_depth
                assert tuple(st[key].shape) == tuple(z[f"s{s_}_{key}"].shape), (s_, key)
        ref_attn, ref_depth = torch.from_numpy(z[f"s{s_}_attn_weight"]), torch.from_numpy(z[f"s{s_}_depth"])
        if s_ > 1:
            drift_free = F.interpolate(drift_free.float()[:, None], scale_factor=2, mode="bilinear", align_corners=True)[:, 0] > 0.999
        agree = (st["depth"] - ref_depth).abs() <= 1e-4 * ref_depth.abs()  # linear sampling can go negative
        stable = top2_gap(ref_attn) > 1e-3
        bad = ((~agree) & stable & drift_free).float().sum().item() / max(1.0, (stable & drift_free).float().sum().item())
        assert bad < 5e-3, f"stage {s_}: {bad:.3%} of tie-free, drift-free pixels differ by > 1e-4 relative"
        if s_ == 1:
            assert (st["attn_weight"] - ref_attn).abs().max().item() < 5e-5
        drift_free = drift_free & agree
    assert out["depth"].data_ptr() == out["stage4"]["depth"].data_ptr()
    conf = torch.from_numpy(z["s4_photometric_confidence"])
    assert (out["photometric_confidence"] - conf).abs()[drift_free].max().item() < 1e-3  # max probability of the last stage, after four stages of fp32 noise
    if GOLDEN_CASES[name].get("mono") and GOLDEN_CASES[name].get("group_cor"):
        # mono_feat = the reference view's features in the NATURAL channel order although the engine keeps stages 2-4 group-interleaved
        assert any(eng.interleave), eng.interleave
        want = oracle.fpn4_features(m.state_dict(), imgs[0])
        for s_ in range(1, 5):
            got = out[f"stage{s_}"]["mono_feat"]
            assert tuple(got.shape) == tuple(want[f"stage{s_}"].shape)
            assert (got - want[f"stage{s_}"]).abs().max().item() <= 2e-5 * want[f"stage{s_}"].abs().max().item(), s_



def test_scan_to_point_cloud_on_cpu(emu, monkeypatch, tmp_path):
    """The evaluation flow around the forward (tests/flow_util.py): scan directory -> prefetcher -> engine -> PFM -> filter ->
    PLY, with the kernels on the emulation library."""
    import contextlib
    from flow_util import scan_to_point_cloud
    from mvster_b200.engine import InferenceEngine
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    m = build_model(SHIPPED, seed=2)
    m.reg_precision = m.fpn_precision = "fp32"
    m.overlap_stages, m.fpn_backend = False, "native"
    eng = InferenceEngine(torch.device("cpu"))
    eng.refresh_weights(m)
    with torch.no_grad():
        scan_to_point_cloud(tmp_path, lambda s: eng.forward(m, s["imgs"], s["proj_matrices"], s["depth_values"]), "cpu", hw=(70, 136))
