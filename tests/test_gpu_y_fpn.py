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


def test_fp16_term_arithmetic_detects_out_of_range_activations_and_falls_back():
    """the default `2xfp16` arithmetic saturates layer inputs at +-65504.  The converters report such inputs
    during the first forward of a weight set; the engine must then warn, switch to three bf16 terms (full fp32 range) and return
    what the exact-fp32 engine returns - not a silently clamped result."""
    import warnings
    from util import SHIPPED, build_model, top2_gap
    from mvster_b200 import synth
    imgs, proj, dv = synth.make_inputs(1, 3, 128, 192, seed=11)
    imgs = [t * 3.0e5 for t in imgs]  # pyramid activations far above the fp16 range
    args = ([t.to(DEV) for t in imgs], {k: v.to(DEV) for k, v in proj.items()}, dv.to(DEV))
    m = build_model(SHIPPED, 4).to(DEV)
    m.use_cuda_graph = False
    with torch.no_grad(), warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        got = m(*args)
    assert any("two-fp16-term" in str(w.message) for w in caught), [str(w.message) for w in caught]
    eng = m._engines[DEV.index]
    assert eng._precision(m, "fpn") == "3xbf16" and eng._precision(m, "reg") == "3xbf16"
    ref_m = build_model(SHIPPED, 4).to(DEV)
    ref_m.use_cuda_graph, ref_m.reg_precision, ref_m.fpn_precision = False, "fp32", "fp32"
    with torch.no_grad():
        want = ref_m(*args)
    stable = top2_gap(want["stage1"]["attn_weight"]) > 1e-3
    bad = ((got["stage1"]["depth"] != want["stage1"]["depth"]) & stable).float().mean().item()
    assert torch.isfinite(got["depth"]).all() and bad < 1e-3, bad
    # in-range inputs: no warning, the fast arithmetic stays
    m2 = build_model(SHIPPED, 4).to(DEV)
    m2.use_cuda_graph = False
    imgs2, proj2, dv2 = synth.make_inputs(1, 3, 128, 192, seed=11)
    with torch.no_grad(), warnings.catch_warnings(record=True) as caught2:
        warnings.simplefilter("always")
        m2([t.to(DEV) for t in imgs2], {k: v.to(DEV) for k, v in proj2.items()}, dv2.to(DEV))
    assert not any("two-fp16-term" in str(w.message) for w in caught2)
    assert m2._engines[DEV.index]._precision(m2, "reg") == "2xfp16"
