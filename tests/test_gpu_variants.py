"""GPU parity of the non-default branches of the hot path: per-channel
squared-difference cost (group_cor=False), attn_fuse_d=False, linear depth sampling, reg3d."""
import numpy as np
import pytest
import torch

from util import GOLDEN_CASES, build_model, load_golden, oracle, oracle_cfg, top2_gap
from test_gpu_parity import DEV, et_inputs, from_ndhwc, ndhwc, nhwc, record

from mvster_b200 import capi, packing

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("group_cor,fuse_d,case", [
    (False, True, (1, 3, 64, 64, 8, 16, 24, 2.0)),   # C = 64 cost channels
    (False, False, (2, 3, 16, 16, 4, 32, 40, 3.0)),  # the reference's `plain` configuration
    (True, False, (1, 4, 32, 8, 8, 24, 32, 1.0)),
    (False, True, (1, 2, 8, 8, 4, 32, 32, 5.0)),
])
def test_et_variants_match_oracle(group_cor, fuse_d, case):
    B, nv, C, G, D, H, W, step = case
    feats, cams, hypo = et_inputs(*case, seed=11)
    want = oracle.et_aggregate(feats, cams, hypo, group_cor, G, 2.0, attn_fuse_d=fuse_d)
    got = capi.et_fuse(nhwc(feats[0]), [nhwc(f) for f in feats[1:]], capi.pose(cams.to(DEV)), hypo.to(DEV), G, 2.0,
                       group_cor=group_cor, fuse_d=fuse_d)
    assert tuple(got.shape) == (B, D, H, W, G)
    err = (from_ndhwc(got) - want).abs().max().item() / want.abs().max().item()
    record(f"et_variant_gc{group_cor}_fd{fuse_d}_{case}", rel_to_max=err)
    assert err < 2e-4


@pytest.mark.parametrize("B,D,H,W", [(1, 8, 8, 16), (2, 4, 32, 48)])
def test_linear_hypotheses(B, D, H, W):
    dv = torch.tensor([[425.0, 600.0, 935.0], [300.0, 500.0, 1200.5]])[:B].contiguous()
    got = capi.hypo_init_linear(dv.to(DEV), D, H, W).cpu()
    want = oracle.hypo_init_linear(dv, D, H, W)
    assert ((got - want).abs() / want.abs()).max().item() < 3e-7
    rng = np.random.RandomState(2)
    depth = torch.from_numpy(rng.uniform(450, 900, (B, H // 2, W // 2)).astype(np.float32))
    itv = 0.5 * (dv[:, -1] - dv[:, 0]) / dv.size(1)
    got = capi.hypo_schedule_linear(depth.to(DEV), dv.to(DEV), 0.5, D, H, W).cpu()
    want = oracle.hypo_schedule_linear(depth, D, itv, H, W)
    rel = ((got - want).abs() / want.abs()).max().item()
    record(f"hypo_schedule_linear_{B}x{D}x{H}x{W}", rel_err=rel)
    assert rel < 1e-6


@pytest.mark.parametrize("k,B,D,H,W", [(0, 1, 8, 16, 24), (3, 2, 4, 32, 32)])
def test_reg3d_matches_oracle(k, B, D, H, W):
    kw = GOLDEN_CASES["reg3d_b1_v2_64x64"]
    sd = build_model(kw, 9).state_dict()
    G, down = kw["group_cor_dim"][k], (3, 3, 2, 2)[k]
    rng = np.random.RandomState(k + H)
    cost = torch.from_numpy((rng.randn(B, G, D, H, W) * 0.05).astype(np.float32))
    want = oracle.reg3d_logits(sd, f"reg.{k}", cost, down)
    blob = packing.pack_reg3d(sd, f"reg.{k}", capi.reg3d_layer_table(G, down))
    got = capi.reg3d(blob.to(DEV), ndhwc(cost), down).cpu()
    err = (got - want).abs().max().item() / want.abs().max().item()
    record(f"reg3d_stage{k + 1}_{B}x{D}x{H}x{W}", rel_to_max=err)
    assert err < 2e-5


@pytest.mark.parametrize("name", ["plain_b1_v2_64x64", "reg3d_b1_v2_64x64"])
def test_variant_teacher_forced_stages(name):
    """Per stage, the oracle's features and hypotheses go through the CUDA kernels of the variant configuration;
    the cost volume and the attention are compared with the oracle's (locates a deviation to one operator)."""
    kw = GOLDEN_CASES[name]
    z, imgs, proj, dv = load_golden(name)
    m = build_model(kw, int(z["meta_seed"]))
    sd, cfg = m.state_dict(), oracle_cfg(kw)
    with torch.no_grad():
        feats = [oracle.fpn4_features(sd, im) for im in imgs]
        ref_out = oracle.cascade_forward(sd, cfg, imgs, proj, dv, features=feats)
    failures = []
    for k in range(4):
        key = f"stage{k + 1}"
        C = feats[0][key].shape[1]
        G = cfg["group_cor_dim"][k] if cfg["group_cor"] else C
        hypo = ref_out[key]["hypo_depth"]
        f = [nhwc(ft[key]) for ft in feats]
        cost = capi.et_fuse(f[0], f[1:], capi.pose(proj[key].to(DEV)), hypo.to(DEV), G, cfg["attn_temp"],
                            group_cor=cfg["group_cor"], fuse_d=cfg["attn_fuse_d"])
        want_cost = ref_out[key]["cost"]
        cerr = (from_ndhwc(cost) - want_cost).abs().max().item() / want_cost.abs().max().item()
        cost_in = ndhwc(want_cost)  # the oracle's cost feeds the regulariser: isolates it from the warp
        if cfg["reg_net"] == "reg2d":
            packed = packing.pack_reg2d(sd, f"reg.{k}", capi.reg2d_layer_table(G))
            feat8 = capi.reg2d(packed["blob"].to(DEV), cost_in)
            h = capi.head(hypo.to(DEV), cfg["depth_interals_ratio"][k], feat8=feat8, prob_w=packed["prob_w"].to(DEV),
                          prob_b=packed["prob_b"].to(DEV), inverse=cfg["inverse_depth"])
        else:
            down = (3, 3, 2, 2)[k]
            blob = packing.pack_reg3d(sd, f"reg.{k}", capi.reg3d_layer_table(G, down))
            h = capi.head(hypo.to(DEV), cfg["depth_interals_ratio"][k], logits=capi.reg3d(blob.to(DEV), cost_in, down), inverse=cfg["inverse_depth"])
        aerr = (h["attn_weight"].cpu() - ref_out[key]["attn_weight"]).abs().max().item()
        stable = top2_gap(ref_out[key]["attn_weight"]) > 1e-3
        d, rd = h["depth"].cpu(), ref_out[key]["depth"]
        bad = (((d - rd).abs() > 1e-4 * rd.abs()) & stable).float().mean().item()
        record(f"variant_teacher_forced_{name}_{key}", cost_vs_oracle=cerr, attn_vs_oracle_given_oracle_cost=aerr, depth_bad_stable=bad,
               stable_frac=stable.float().mean().item(), hypo_min=float(hypo.min()), hypo_max=float(hypo.max()))
        if cerr > 2e-4:
            failures.append(f"{key}: cost off by {cerr:.2e} of max")
        if aerr > 1e-4:
            failures.append(f"{key}: attention off by {aerr:.2e} given the oracle's cost volume")
        if bad != 0.0:
            failures.append(f"{key}: {bad:.3%} tie-free pixels disagree on depth")
    assert not failures, "; ".join(failures)


@pytest.mark.parametrize("name", ["reg3d_b1_v2_64x64", "plain_b1_v2_64x64"])
def test_variant_module_forward_against_reference_golden(name):
    """MVS4net.forward on the GPU for the reg3d and the non-group / linear-depth / attn_fuse_d=False
    configurations vs outputs of the unmodified reference."""
    z, imgs, proj, dv = load_golden(name)
    torch.backends.cudnn.allow_tf32 = False
    m = build_model(GOLDEN_CASES[name], int(z["meta_seed"])).to(DEV)
    with torch.no_grad():
        out = m([t.to(DEV) for t in imgs], {k: v.to(DEV) for k, v in proj.items()}, dv.to(DEV))
    ok = torch.ones_like(torch.from_numpy(z["s1_depth"]), dtype=torch.bool)
    for s in range(1, 5):
        st = out[f"stage{s}"]
        ref_attn, ref_depth = torch.from_numpy(z[f"s{s}_attn_weight"]), torch.from_numpy(z[f"s{s}_depth"])
        assert tuple(st["attn_weight"].shape) == tuple(ref_attn.shape)
        assert ("inverse_min_depth" in st) == (f"s{s}_inverse_min_depth" in z.files)
        if s > 1:
            ok = torch.nn.functional.interpolate(ok.float()[:, None], scale_factor=2, mode="bilinear", align_corners=True)[:, 0] > 0.999
        agree = (st["depth"].cpu() - ref_depth).abs() <= 1e-4 * ref_depth.abs()  # linear sampling can go negative
        stable = top2_gap(ref_attn) > 1e-3
        bad = ((~agree) & stable & ok).float().sum().item() / max(1.0, (stable & ok).float().sum().item())
        record(f"e2e_golden_{name}_s{s}", bad_frac=bad, considered=float((stable & ok).float().mean()), agree_all=float(agree.float().mean()))
        assert bad < 5e-3, f"stage {s}: {bad:.3%} of tie-free, drift-free pixels differ"
        if s == 1:
            assert (st["attn_weight"].cpu() - ref_attn).abs().max().item() < 5e-4  # default engine: 3xTF32 feature pyramid
        ok = ok & agree
