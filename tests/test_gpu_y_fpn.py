"""FPN4 on libmvster_b200 (mvster_b200/fpn_engine.py) vs the oracle's FPN (mvs4net_utils.py:472-502), and the
whole forward with the native feature pyramid vs the reference's golden outputs."""
import pytest
import torch
import torch.nn.functional as F

from util import GOLDEN_CASES, SHIPPED, build_model, load_golden, oracle, top2_gap
from test_gpu_parity import DEV, record

from mvster_b200 import fpn_engine, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fused_last", [True, False], ids=["fused_last_level", "literal_last_level"])
@pytest.mark.parametrize("npass,gen,tol", [(0, 2, 2e-5), (3, 2, 1e-4), (1, 2, 2e-2), (3, 3, 2e-5), (2, 3, 2e-5)])  # (2, 3): two fp16 terms
@pytest.mark.parametrize("N,H,W", [(2, 64, 128), (3, 128, 192)])
def test_native_fpn_matches_oracle(npass, gen, tol, N, H, W, fused_last):
    sd = build_model(SHIPPED, 4).state_dict()
    imgs, _, _ = synth.make_inputs(1, N, H, W, seed=8)
    x = torch.cat(imgs, 0)
    with torch.no_grad():
        want = oracle.fpn4_features(sd, x)
    wts = {k: v.to(DEV) for k, v in fpn_engine.pack_fpn(sd).items()}
    got = fpn_engine.run_fpn(wts, x.to(DEV), npass, fused_last=fused_last, gen=gen)
    for s in range(1, 5):
        g = got[f"stage{s}"].permute(0, 3, 1, 2).cpu()
        w = want[f"stage{s}"]
        assert g.shape == w.shape
        err = (g - w).abs().max().item() / w.abs().max().item()
        record(f"fpn_npass{npass}_gen{gen}_{'fused' if fused_last else 'literal'}_{N}x{H}x{W}_stage{s}", rel_to_max=err)
        assert err < tol, f"stage{s}: {err:.2e}"


@pytest.mark.parametrize("backend", ["torch", "native"])
def test_cuda_graph_replay_matches_eager(backend):
    """forward_graphed (one CUDA graph per input signature) must reproduce the eager launch sequence bit for bit,
    also when replayed on new input values."""
    m = build_model(SHIPPED, 2).to(DEV)
    m.fpn_backend = backend
    torch.backends.cudnn.allow_tf32 = False
    outs = {}
    for graph in (False, True):
        m.use_cuda_graph = graph
        for seed in (5, 6, 5):
            imgs, proj, dv = synth.make_inputs(1, 3, 64, 128, seed=seed)
            with torch.no_grad():
                o = m([t.to(DEV) for t in imgs], {k: v.to(DEV) for k, v in proj.items()}, dv.to(DEV))
            outs[(graph, seed)] = {k: o[k].clone() for k in ("depth", "attn_weight", "photometric_confidence")}
    for seed in (5, 6):
        for k in ("depth", "attn_weight", "photometric_confidence"):
            assert torch.equal(outs[(True, seed)][k], outs[(False, seed)][k]), (seed, k)


@pytest.mark.parametrize("precision", ["fp32", "3xtf32", "3xbf16", "2xfp16"])
def test_forward_with_native_fpn_against_reference_golden(precision):
    name = "shipped_b1_v3_64x128"
    z, imgs, proj, dv = load_golden(name)
    m = build_model(GOLDEN_CASES[name], int(z["meta_seed"])).to(DEV)
    m.fpn_backend, m.fpn_precision = "native", precision
    m.reg_precision = precision
    with torch.no_grad():
        out = m([t.to(DEV) for t in imgs], {k: v.to(DEV) for k, v in proj.items()}, dv.to(DEV))
    ok = torch.ones_like(torch.from_numpy(z["s1_depth"]), dtype=torch.bool)
    for s in range(1, 5):
        st = out[f"stage{s}"]
        ref_attn, ref_depth = torch.from_numpy(z[f"s{s}_attn_weight"]), torch.from_numpy(z[f"s{s}_depth"])
        if s > 1:
            ok = F.interpolate(ok.float()[:, None], scale_factor=2, mode="bilinear", align_corners=True)[:, 0] > 0.999
        agree = (st["depth"].cpu() - ref_depth).abs() <= 1e-4 * ref_depth
        stable = top2_gap(ref_attn) > 1e-3
        bad = ((~agree) & stable & ok).float().sum().item() / max(1.0, (stable & ok).float().sum().item())
        record(f"e2e_native_fpn_{precision}_s{s}", bad_frac=bad, agree_all=float(agree.float().mean()),
               attn_err_s1=float((st["attn_weight"].cpu() - ref_attn).abs().max()) if s == 1 else -1.0)
        assert bad < 5e-3, f"stage {s}: {bad:.3%} of tie-free, drift-free pixels differ"
        ok = ok & agree
