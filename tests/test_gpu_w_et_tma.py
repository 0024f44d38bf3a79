"""TMA-staged warp + ET kernel (csrc/et_fuse_tma.cuh) on the GPU: against the CPU oracle, against the global-gather window
kernel, in partial / accumulate mode, on ragged multi-tile frames, with hypotheses that leave the staged boxes (fallback path),
with both tile heights, and at the cfg2 stage-3 / stage-4 sizes.  Tolerance as for the other warp/ET kernels: <= 2e-4 of max.
reference: models/mvs4net_utils.py:13-59, :1037-1060."""
import numpy as np
import pytest
import torch

from util import narrow_et_inputs, oracle
from test_gpu_parity import DEV, from_ndhwc, nhwc, record

from mvster_b200 import capi

pytestmark = pytest.mark.gpu

TMA_CASES = [  # (B, nv, C, G, D, H, W, step_deg, rel_span)
    (1, 6, 8, 4, 4, 37, 70, 1.0, 0.12),
    (2, 3, 8, 4, 4, 64, 160, 2.0, 0.3),
    (1, 5, 16, 4, 4, 33, 96, 1.5, 0.2),
    (2, 4, 32, 8, 8, 24, 72, 3.0, 0.06),
    (1, 13, 8, 4, 4, 32, 64, 0.5, 0.1),    # 12 source views: every tensor-map slot of a launch
    (1, 5, 8, 4, 4, 128, 320, 1.0, 0.05),  # many tiles per CTA
]


@pytest.mark.parametrize("case", TMA_CASES)
def test_tma_kernel_matches_oracle_and_window_kernel(case, monkeypatch):
    B, nv, C_, G, D, H, W, step, span = case
    feats, cams, hypo = narrow_et_inputs(B, nv, C_, D, H, W, step, span, seed=17)
    hypo[:, :, 0, :3] *= torch.tensor([1.0, 0.8, 0.6, 0.4] * (D // 4)).reshape(1, D, 1)  # these pixels leave the staged boxes
    want = oracle.et_aggregate(feats, cams, hypo, True, G, 2.0)
    scale = want.abs().max().item()
    ref, srcs, pose, hy = nhwc(feats[0]), [nhwc(f) for f in feats[1:]], capi.pose(cams.to(DEV)), hypo.to(DEV)
    got = from_ndhwc(capi.et_fuse(ref, srcs, pose, hy, G, 2.0, tma=True))
    assert "tma" in capi.et_last_kernel(), capi.et_last_kernel()
    err = (got - want).abs().max().item() / scale
    win = from_ndhwc(capi.et_fuse(ref, srcs, pose, hy, G, 2.0, tma=False))
    assert "tma" not in capi.et_last_kernel()
    dev = (got - win).abs().max().item() / scale
    record(f"et_tma_{C_}_{H}x{W}_v{nv - 1}", vs_oracle=err, vs_window_kernel=dev)
    assert torch.isfinite(got).all() and err <= 2e-4 and dev <= 3e-5
    # group-interleaved channel layout (MVSTER_ET_INTERLEAVED): same features with permuted channels, both window kernels
    perm = capi.interleave_perm(C_, G)
    ref_il, srcs_il = ref[..., perm].contiguous(), [s_[..., perm].contiguous() for s_ in srcs]
    for tma in (True, False):
        got_il = from_ndhwc(capi.et_fuse(ref_il, srcs_il, pose, hy, G, 2.0, tma=tma, interleaved=True))
        assert "interleaved" in capi.et_last_kernel(), capi.et_last_kernel()
        assert (got_il - got).abs().max().item() <= 1e-5 * scale, tma
    k = max(1, (nv - 1) // 2)
    wsum = torch.empty(B, D, H, W, device=DEV)
    cost = capi.et_fuse(ref, srcs[:k], pose[:, :k].contiguous(), hy, G, 2.0, wsum=wsum, partial=True, tma=True)
    if k < nv - 1:
        capi.et_fuse(ref, srcs[k:], pose[:, k:].contiguous(), hy, G, 2.0, cost=cost, wsum=wsum, partial=True, accumulate=True, tma=True)
    capi.et_normalize(cost, wsum)
    assert (from_ndhwc(cost) - want).abs().max().item() <= 2e-4 * scale
    if C_ == 8:
        monkeypatch.setenv("MVSTER_ET_TMA_TH", "15")
        got15 = from_ndhwc(capi.et_fuse(ref, srcs, pose, hy, G, 2.0, tma=True))
        assert "15x32" in capi.et_last_kernel()
        assert (got15 - want).abs().max().item() <= 2e-4 * scale
    # determinism: same bits on a second launch
    again = from_ndhwc(capi.et_fuse(ref, srcs, pose, hy, G, 2.0, tma=True))
    assert torch.equal(again, got15 if C_ == 8 else got)


@pytest.mark.parametrize("C_,G,D,H,W", [(8, 4, 4, 512, 640), (16, 4, 4, 256, 320), (32, 8, 8, 128, 160)], ids=["stage4", "stage3", "stage2"])
def test_tma_kernel_at_cfg2_stage_sizes(C_, G, D, H, W):
    """cfg2 stage shapes with 4 source views and per-pixel random hypothesis centres: the TMA kernel against the per-hypothesis
    tiled kernel (itself checked against the oracle above and in test_gpu_parity.py) - the oracle would take minutes here."""
    feats, cams, hypo = narrow_et_inputs(1, 5, C_, D, H, W, 1.0, {8: 0.01, 16: 0.02, 32: 0.06}[C_], seed=5)
    ref, srcs, pose, hy = nhwc(feats[0]), [nhwc(f) for f in feats[1:]], capi.pose(cams.to(DEV)), hypo.to(DEV)
    got = capi.et_fuse(ref, srcs, pose, hy, G, 2.0, tma=True)
    assert "tma" in capi.et_last_kernel()
    want = capi.et_fuse(ref, srcs, pose, hy, G, 2.0, window=False)
    err = (got - want).abs().max().item() / want.abs().max().item()
    record(f"et_tma_cfg2_C{C_}", vs_tiled_kernel=err)
    assert torch.isfinite(got).all() and err <= 5e-5
