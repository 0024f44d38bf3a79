"""The other BASELINE.json configurations as parity cases: cfg3's view count (9 source views) end to end against the oracle,
and the large-frame shapes of cfg2' / cfg4 (DTU raw 1152x1536 as the reference's loader crops it; 1024x1920) through
size-independent properties plus agreement between the tensor-core engine and the exact-fp32 CUDA-core engine on the same
frame.  Frame sizes are multiples of 64: the 4-stage cascade runs reg2d (three stride-2 levels) at 1/8 resolution, and the
reference's own skip additions fail otherwise (cfg0's 1 ref + 1 src case is the golden fixture shipped_b2_v2_64x64)."""
import pytest
import torch

from util import SHIPPED, build_model, oracle, oracle_cfg, top2_gap
from test_gpu_parity import DEV, record

from mvster_b200 import synth

pytestmark = pytest.mark.gpu


def _forward(m, imgs, proj, dv):
    with torch.no_grad():
        out = m([t.to(DEV) for t in imgs], {k: v.to(DEV) for k, v in proj.items()}, dv.to(DEV))
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("nv,H,W", [(10, 128, 192)], ids=["cfg3_9src_128x192"])
def test_view_counts_against_oracle(nv, H, W):
    """End to end against the CPU oracle (tie-aware, drift-aware, as for the golden fixtures)."""
    imgs, proj, dv = synth.make_inputs(1, nv, H, W, seed=31)
    m = build_model(SHIPPED, 13)
    with torch.no_grad():
        ref = oracle.cascade_forward(m.state_dict(), oracle_cfg(SHIPPED), imgs, proj, dv)
    out = _forward(m.to(DEV), imgs, proj, dv)
    ok = torch.ones_like(ref["stage1"]["depth"], dtype=torch.bool)
    for s in range(1, 5):
        st, rs = out[f"stage{s}"], ref[f"stage{s}"]
        if s > 1:
            ok = torch.nn.functional.interpolate(ok.float()[:, None], scale_factor=2, mode="bilinear", align_corners=True)[:, 0] > 0.999
        agree = (st["depth"].cpu() - rs["depth"]).abs() <= 1e-4 * rs["depth"].abs()
        stable = top2_gap(rs["attn_weight"]) > 1e-3
        bad = ((~agree) & stable & ok).float().sum().item() / max(1.0, (stable & ok).float().sum().item())
        record(f"config_nv{nv}_{H}x{W}_s{s}", bad_frac=bad, agree_all=float(agree.float().mean()))
        if s == 1:
            aerr = (st["attn_weight"].cpu() - rs["attn_weight"]).abs().max().item()
            assert aerr < 2e-3, f"stage-1 probabilities differ by {aerr:.2e}"
        assert bad < 5e-3, f"stage {s}: {bad:.3%} of tie-free, drift-free pixels differ by > 1e-4 relative"
        ok = ok & agree


@pytest.mark.parametrize("nv,H,W", [(5, 1152, 1536), (3, 1024, 1920)], ids=["dtu_raw_1152x1536", "tnt_1024x1920"])
def test_large_frames_properties_and_engine_agreement(nv, H, W):
    imgs, proj, dv = synth.make_inputs(1, nv, H, W, seed=5)
    m = build_model(SHIPPED, 3).to(DEV)
    out = _forward(m, imgs, proj, dv)                       # default engine: 3 x bf16 tensor-core convolutions, native FPN
    for s in range(1, 5):
        st = out[f"stage{s}"]
        attn, hypo, depth = st["attn_weight"], st["hypo_depth"], st["depth"]
        assert torch.isfinite(attn).all() and torch.isfinite(depth).all()
        assert (attn.sum(1) - 1).abs().max().item() < 1e-5
        assert (depth[:, None] == hypo).any(1).all()
        assert (hypo[:, :-1] > hypo[:, 1:]).all()
    keep = {s: (out[f"stage{s}"]["attn_weight"].clone(), out[f"stage{s}"]["depth"].clone()) for s in (1, 4)}
    m.reg_precision, m.fpn_precision = "fp32", "fp32"       # every convolution as an exact fp32 FMA chain on the CUDA cores
    ref = _forward(m, imgs, proj, dv)
    a_tc, _ = keep[1]
    a_err = (a_tc - ref["stage1"]["attn_weight"]).abs().max().item()
    stable = top2_gap(ref["stage1"]["attn_weight"]) > 1e-3
    bad = ((out["stage1"]["depth"] != ref["stage1"]["depth"]) & stable).float().mean().item()
    record(f"large_{H}x{W}_tc_vs_fp32", stage1_attn_abs=a_err, stage1_depth_bad=bad,
           stage4_depth_equal=float((keep[4][1] == ref["stage4"]["depth"]).float().mean()))
    # max over ~2 M softmax outputs of two differently rounded U-Nets; the winner-take-all depth is the criterion that matters
    assert a_err < 1e-2 and bad < 1e-3, (a_err, bad)
