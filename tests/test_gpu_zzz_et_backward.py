"""Backward of the fused warp + epipolar-Transformer kernel on the GPU (mvster_et_fuse_bwd_f32 through the autograd node of
mvster_b200/train_ops.py): against gradients taken through the unmodified reference (tests/golden/et_backward.npz), against
fp64 autograd through the oracle's formulation at stage-like shapes, and inside a training step of ``MVS4net``."""
import numpy as np
import pytest
import torch

from util import SHIPPED, build_model, narrow_et_inputs, oracle
from test_et_backward_golden import CASES, load_case

from mvster_b200 import _lib, synth, train_ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("name", CASES)
def test_backward_kernel_matches_reference_gradients(name):
    c = load_case(name)
    leaves = [f.to(DEV).requires_grad_(True) for f in c["feats"]]
    n0 = _lib.launch_count()
    cost = train_ops.aggregate(leaves, c["cams"].to(DEV), c["hypo"].to(DEV), c["G"], 2.0)
    assert (cost.detach().cpu() - c["cost"]).abs().max().item() <= 2e-4 * c["cost"].abs().max().item()
    grads = torch.autograd.grad(cost, leaves, c["gout"].to(DEV))
    assert 4 <= _lib.launch_count() - n0 <= 5                    # pose, [tile boxes,] forward (partial), normalise, backward
    for v in range(c["nv"]):
        scale = c["grads"][v].abs().max().item()
        err = (grads[v].cpu() - c["grads"][v]).abs().max().item()
        assert err <= 2e-4 * scale, (v, err, scale)             # CPU emulation of the same source: <= 2.1e-6


@pytest.mark.parametrize("shape", [(1, 5, 8, 4, 4, 64, 80, 1.0, 0.02), (1, 3, 16, 4, 4, 32, 48, 2.0, 0.05),
                                   (2, 3, 32, 8, 8, 16, 24, 1.0, 0.1), (1, 4, 64, 8, 8, 16, 20, 1.0, 0.5)])
def test_backward_kernel_matches_fp64_autograd(shape):
    """Stage-like hypothesis ranges (many pixels share taps -> colliding atomics) against fp64 autograd on the CPU."""
    B, nv, C, G, D, H, W, step, span = shape
    feats, cams, hypo = narrow_et_inputs(B, nv, C, D, H, W, step, span, seed=31)
    gout = torch.from_numpy(np.random.RandomState(3).randn(B, G, D, H, W).astype(np.float32))
    f64 = [f.double().requires_grad_(True) for f in feats]
    want = torch.autograd.grad(oracle.et_aggregate(f64, cams.double(), hypo.double(), True, G, 2.0), f64, gout.double())
    leaves = [f.to(DEV).requires_grad_(True) for f in feats]
    got = torch.autograd.grad(train_ops.aggregate(leaves, cams.to(DEV), hypo.to(DEV), G, 2.0), leaves, gout.to(DEV))
    for v in range(nv):
        scale = want[v].abs().max().item()
        assert (got[v].cpu() - want[v].float()).abs().max().item() <= 2e-4 * scale


def test_frozen_views_are_skipped_and_repeat_runs_agree():
    feats, cams, hypo = narrow_et_inputs(1, 3, 8, 4, 64, 80, 1.0, 0.02, seed=5)
    gout = torch.ones(1, 4, 4, 64, 80, device=DEV)
    full = [f.to(DEV).requires_grad_(True) for f in feats]
    g_full = torch.autograd.grad(train_ops.aggregate(full, cams.to(DEV), hypo.to(DEV), 4, 2.0), full, gout)
    some = [f.to(DEV).requires_grad_(i != 1) for i, f in enumerate(feats)]
    g_some = torch.autograd.grad(train_ops.aggregate(some, cams.to(DEV), hypo.to(DEV), 4, 2.0), [some[0], some[2]], gout)
    assert torch.equal(g_some[0], g_full[0])                                        # no atomics on the reference gradient
    assert torch.allclose(g_some[1], g_full[2], rtol=0, atol=1e-5 * g_full[2].abs().max().item())   # atomic order only


def test_training_step_with_kernels_matches_pytorch_ops():
    """MVS4net in train mode: loss gradients w.r.t. every parameter with the aggregation on the kernels vs on PyTorch ops."""
    from mvster_b200 import MVS4net_loss
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = build_model(SHIPPED, seed=3).to(DEV).train()
    for m in model.modules():
        if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
            m.eval()
    imgs, proj, dv = synth.make_inputs(1, 3, 64, 64, seed=4, step_deg=1.0)
    imgs = [i.to(DEV) for i in imgs]
    proj = {k: v.to(DEV) for k, v in proj.items()}
    dv = dv.to(DEV)
    gt = {f"stage{k + 1}": torch.full((1, 64 >> (3 - k), 64 >> (3 - k)), 680.0, device=DEV) for k in range(4)}
    mask = {k: torch.ones_like(v) for k, v in gt.items()}
    grads = {}
    for use in (False, True):
        model.stagenet.train_et = use
        model.zero_grad(set_to_none=True)
        out = model(imgs, proj, dv)
        loss = MVS4net_loss(out, gt, mask, stage_lw=[1, 1, 1, 1], l1ot_lw=[1, 1], inverse_depth=True, mono=True)[0]
        loss.backward()
        grads[use] = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    assert grads[True].keys() == grads[False].keys() and len(grads[True]) > 50
    # the two forwards differ by fp32 rounding (attn_weight up to 2e-3 at stage 4 with these random weights, identical argmax),
    # which bounds how closely the gradients can agree; prob.bias has a zero true gradient (softmax shift invariance) - skipped
    top = max(g.abs().max().item() for g in grads[False].values())
    worst = max(((grads[True][n] - grads[False][n]).abs().max() / grads[False][n].abs().max()).item()
                for n in grads[True] if grads[False][n].abs().max().item() > 1e-5 * top)
    assert worst < 5e-2, worst                                      # measured on the CPU emulation: 7.6e-3


def test_warp_edge_cases_of_the_reference_through_the_kernels():
    """Forward kernels on the reference's own homo_warping edge-case fixture (out of bounds on every side, z == 0, source size
    != reference size); see util.warp_edge_through_et.  The CPU emulation of the same sources agrees to 2e-6 of max."""
    from util import warp_edge_through_et
    from mvster_b200 import capi
    outs, want = warp_edge_through_et(capi, DEV)
    for got in outs:
        assert torch.isfinite(got).all()
        assert (got - want).abs().max().item() <= 1e-5 * want.abs().max().item()
        assert ((got == 0) == (want == 0)).float().mean() > 0.999


def test_more_source_views_than_one_launch_takes():
    """19 source views > MVSTER_MAX_VIEWS: chained forward launches and chunked backward against the oracle."""
    from mvster_b200 import capi
    rng = np.random.RandomState(17)
    B, nv, C, G, D, H, W = 1, 20, 8, 4, 4, 16, 24
    feats = [torch.from_numpy(rng.randn(B, C, H, W).astype(np.float32)) for _ in range(nv)]
    cams = synth.stage_projections(synth.arc_cameras(nv, H, W, 0.3), B, num_stage=1)["stage1"]
    hypo = oracle.hypo_init_inverse(torch.tensor([[425.0, 935.0]] * B), D, H, W)
    want = oracle.et_aggregate(feats, cams, hypo, True, G, 2.0)
    nhwc = [f.permute(0, 2, 3, 1).contiguous().to(DEV) for f in feats]
    got = capi.et_fuse(nhwc[0], nhwc[1:], capi.pose(cams.to(DEV)), hypo.to(DEV), G, 2.0).permute(0, 4, 1, 2, 3).cpu()
    assert (got - want).abs().max().item() <= 2e-4 * want.abs().max().item()
    gout = torch.from_numpy(rng.randn(B, G, D, H, W).astype(np.float32))
    f64 = [f.double().requires_grad_(True) for f in feats]
    gwant = torch.autograd.grad(oracle.et_aggregate(f64, cams.double(), hypo.double(), True, G, 2.0), f64, gout.double())
    leaves = [f.to(DEV).requires_grad_(True) for f in feats]
    ggot = torch.autograd.grad(train_ops.aggregate(leaves, cams.to(DEV), hypo.to(DEV), G, 2.0), leaves, gout.to(DEV))
    for v in range(nv):
        assert (ggot[v].cpu() - gwant[v].float()).abs().max().item() <= 2e-4 * gwant[v].abs().max().item(), v
