"""CPU-side tests: module mirror (state_dict, autograd path, losses), weight folding/packing and the
C-ABI surface (library loads, exports every symbol declared in include/mvster_b200.h).  No GPU compute."""
import ctypes
import re

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from util import GOLDEN, GOLDEN_CASES, REPO, SHIPPED, build_model, load_golden, oracle

from mvster_b200 import Blend_loss, MVS4net, MVS4net_loss, _lib, packing


def test_header_symbols_exported_and_bound():
    header = (REPO / "include" / "mvster_b200.h").read_text()
    declared = set(re.findall(r"\b(mvster_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    lib = _lib.load()  # raises if the .so is missing: the build is part of the contract
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported"
    assert lib.mvster_version() >= 100
    assert isinstance(lib.mvster_last_error(), bytes)


def test_c_abi_argument_errors_do_not_need_a_gpu():
    lib = _lib.load()
    rc = lib.mvster_reg2d_layer_info(8, 99, (ctypes.c_int64 * 8)())
    assert rc == -1 and b"bad layer" in lib.mvster_last_error()
    info = (ctypes.c_int64 * 8)()
    assert lib.mvster_reg2d_layer_info(8, 6, info) == 0
    assert list(info)[:5] == [64, 64, 3, 1, 0] and info[7] == 27
    # blob = sum over layers of taps*cin*cout + cout
    tbl = [(8, 8, 9), (8, 16, 9), (16, 16, 27), (16, 32, 9), (32, 32, 27), (32, 64, 9), (64, 64, 27), (64, 32, 9), (32, 16, 9), (16, 8, 9)]
    assert lib.mvster_reg2d_blob_floats(8) == sum(t * a * b + b for a, b, t in tbl)
    with pytest.raises(_lib.MvsterLibraryError):
        _lib.check(-1, "probe")


def test_inference_on_cpu_raises_instead_of_falling_back():
    m = build_model(SHIPPED, 1)
    z, imgs, proj, dv = load_golden("shipped_b2_v2_64x64")
    with torch.no_grad(), pytest.raises(_lib.MvsterLibraryError):
        m(imgs, proj, dv)


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_autograd_path_matches_reference_eval(name):
    """The differentiable PyTorch formulation (training path), run in eval mode with grad enabled."""
    z, imgs, proj, dv = load_golden(name)
    m = build_model(GOLDEN_CASES[name], int(z["meta_seed"]))
    out = m._forward_autograd(imgs, proj, dv)
    for s in range(1, 5):
        st = out[f"stage{s}"]
        ref_attn = torch.from_numpy(z[f"s{s}_attn_weight"])
        assert (st["attn_weight"] - ref_attn).abs().max().item() < 5e-6
        d, rd = st["depth"].detach(), torch.from_numpy(z[f"s{s}_depth"])
        assert ((d - rd).abs() > 1e-4 * rd.abs()).float().mean().item() < 0.02  # argmax near-ties only
        assert st["photometric_confidence"].shape == tuple(z[f"s{s}_photometric_confidence"].shape)


def test_training_mode_forward_and_backward():
    z = np.load(GOLDEN / "train_mode.npz")
    g, imgs, proj, dv = load_golden("shipped_b1_v3_64x128")
    m = build_model(SHIPPED, int(g["meta_seed"])).train()
    for mod in m.modules():
        if isinstance(mod, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
            mod.eval()
    out = m(imgs[:2], proj_sub(proj, 2), dv)
    for s in range(1, 5):
        assert (out[f"stage{s}"]["attn_weight"].detach() - torch.from_numpy(z[f"s{s}_attn_weight"])).abs().max() < 5e-6
        if s > 1:
            md, rmd = out[f"stage{s}"]["mono_depth"].detach(), torch.from_numpy(z[f"s{s}_mono_depth"])
            assert (md - rmd).abs().max() < 1e-3 * rmd.abs().max()
        assert out[f"stage{s}"]["photometric_confidence"].item() == 0.0
    loss = sum(out[f"stage{s}"]["attn_weight"].pow(2).mean() for s in range(1, 5))
    loss.backward()
    assert m.reg[3].conv0.conv.weight.grad is not None and m.feature.conv0[0].conv.weight.grad.abs().sum() > 0


def proj_sub(proj, nv):
    return {k: v[:, :nv].contiguous() for k, v in proj.items()}


def test_losses_match_reference():
    z = np.load(GOLDEN / "losses.npz")
    g, imgs, proj, dv = load_golden("shipped_b1_v3_64x128")
    stages = {}
    for s in range(1, 5):
        stages[f"stage{s}"] = {k: torch.from_numpy(g[f"s{s}_{k}"]) for k in ("depth", "hypo_depth", "attn_weight")}
    gt = {f"stage{s}": torch.from_numpy(z[f"gt_stage{s}"]) for s in range(1, 5)}
    mask = {f"stage{s}": torch.from_numpy(z[f"mask_stage{s}"]) for s in range(1, 5)}
    for tag, kw in (("disc", dict(inverse_depth=True, ot_iter=3, ot_eps=1, ot_continous=False)),
                    ("cont", dict(inverse_depth=True, ot_iter=5, ot_eps=0.5, ot_continous=True, stage_lw=[1, 2, 3, 4]))):
        total, l1, ot, oor = MVS4net_loss(stages, gt, mask, **kw)
        assert abs(total.item() - float(z[f"{tag}_total"])) <= 1e-5 * abs(float(z[f"{tag}_total"]))
        assert np.allclose([x.item() for x in ot], z[f"{tag}_ot"], rtol=1e-5)
        assert np.allclose([x.item() for x in oor], z[f"{tag}_oor"], rtol=1e-6)
    res = Blend_loss(stages, gt, mask, inverse_depth=True, depth_max=torch.tensor([935.0]), depth_min=torch.tensor([425.0]))
    assert abs(res[0].item() - float(z["blend_total"])) <= 1e-5 * abs(float(z["blend_total"]))
    assert np.allclose([res[4].item(), res[5].item(), res[6].item()], z["blend_metrics"], rtol=1e-5)


def test_bn_folding_and_packing_against_oracle():
    """Folded [tap][Cin][Cout] weights reproduce conv+BN(+ReLU) of the oracle for a conv and a transposed layer."""
    m = build_model(SHIPPED, 3)
    sd = m.state_dict()
    x = torch.randn(1, 16, 3, 6, 8)
    s, t = packing.bn_scale_shift(sd, "reg.0.conv2.bn")
    w, b = packing.fold_conv3d(sd["reg.0.conv2.conv.weight"], s, t)
    assert w.shape == (27, 16, 16)
    w_t = w.reshape(3, 3, 3, 16, 16).permute(4, 3, 0, 1, 2).contiguous()
    mine = F.relu(F.conv3d(x, w_t, b, 1, 1))
    assert torch.allclose(mine, oracle._cbr3(x, sd, "reg.0.conv2"), atol=2e-6)
    x = torch.randn(1, 64, 2, 3, 4)
    s, t = packing.bn_scale_shift(sd, "reg.0.conv7.1")
    w, b = packing.fold_deconv3d(sd["reg.0.conv7.0.weight"], s, t)
    assert w.shape == (9, 64, 32)
    w_t = w.reshape(1, 3, 3, 64, 32).permute(3, 4, 0, 1, 2).contiguous()
    mine = F.relu(F.conv_transpose3d(x, w_t, b, (1, 2, 2), (0, 1, 1), (0, 1, 1)))
    assert torch.allclose(mine, oracle._up3(x, sd, "reg.0.conv7", (1, 2, 2), (0, 1, 1), (0, 1, 1)), atol=2e-6)


def test_pack_reg2d_blob_layout():
    from mvster_b200 import capi
    m = build_model(SHIPPED, 3)
    tbl = capi.reg2d_layer_table(4)
    packed = packing.pack_reg2d(m.state_dict(), "reg.3", tbl)
    assert packed["blob"].numel() == capi.reg2d_blob_floats(4)
    assert packed["prob_w"].shape == (8,) and packed["prob_b"].shape == (1,)
    L = tbl[2]
    w = packed["blob"][L["w_off"]:L["w_off"] + 27 * 16 * 16].reshape(27, 16, 16)
    s, t = packing.bn_scale_shift(m.state_dict(), "reg.3.conv2.bn")
    ref_w, ref_b = packing.fold_conv3d(m.state_dict()["reg.3.conv2.conv.weight"], s, t)
    assert torch.equal(w, ref_w) and torch.equal(packed["blob"][L["b_off"]:L["b_off"] + 16], ref_b)


def test_tc3_weight_slabs_both_arithmetics():
    """packing.pack_tc3_weights: slab = [2 K-halves][w1 | w2 | w3 rows][8 x 16 bit] per MMA slot, in the plan's order.  The bf16
    terms (split=3) must sum to the fp32 weight to 24 bits, the fp16 terms (split=2) as w1 + 2^-11 w2 to 22 bits; both blobs
    have the size the library expects, and the fp16 blob leaves the third row block empty."""
    from mvster_b200 import _lib, capi
    torch.manual_seed(5)
    cin, cout, kd, k = 32, 16, 3, 3
    w = torch.randn(kd * k * k, cin, cout) / 17.0
    lib = _lib.load()
    plan = capi.conv_tc3_plan(cin, kd, k, 1)
    assert len(plan) == kd * k * k * (cin // 16)
    for split in (3, 2):
        blob = packing.pack_tc3_weights(w, kd, k, 1, split)
        assert blob.numel() * 4 == lib.mvster_conv_tc3_packed_bytes(cin, cout, kd, k, 1)
        raw = blob.view(torch.int16).reshape(len(plan), 2, 3 * 16, 8)
        terms = raw.view(torch.bfloat16 if split == 3 else torch.float16).float()
        for i, (kz, ky, kx, c0, _, _) in enumerate(plan):
            want = w[(kz * k + ky) * k + kx, c0:c0 + 16, :]                       # [16 channels][Cout]
            rows = torch.cat([terms[i, 0], terms[i, 1]], dim=1)                    # [3*N rows][16 channels]
            t1, t2, t3 = rows[0:16].T, rows[16:32].T, rows[32:48].T               # each [16 channels][Cout]
            if split == 3:
                got, tol = t1 + t2 + t3, 2.0 ** -23
            else:
                got, tol = t1 + t2 / 2048.0, 2.0 ** -21
                assert not t3.any()
            # absolute floor 2^-35: below fp16's normal range (|w| < 6.1e-5) the first term is a subnormal, the scaled residual still holds 2^-36
            assert ((got.double() - want.double()).abs() <= tol * want.abs().double() + 2.0 ** -35).all(), (split, i)
    t1, t2 = packing.fp16_split2(torch.tensor([1e5, -1e5, 3.0e-6, 0.0]))
    assert torch.isfinite(t1.float()).all() and torch.isfinite(t2.float()).all()   # out-of-range weights saturate, never inf


def test_dropin_models_package_exports_reference_names():
    import importlib
    import sys
    sys.path.insert(0, str(REPO / "dropin"))
    try:
        sys.modules.pop("models", None)
        models = importlib.import_module("models")
        assert models.MVS4net is MVS4net and models.MVS4net_loss is MVS4net_loss and models.Blend_loss is Blend_loss
    finally:
        sys.path.remove(str(REPO / "dropin"))
        sys.modules.pop("models", None)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference algorithm on the host cores) runs without a GPU and prints ONE JSON line with
    the keys the driver reads; under torchrun only rank 0 prints."""
    import json
    import os
    import subprocess
    import sys
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "depth-maps/s" and j["higher_is_better"] is True and j["value"] > 0
    assert j["metric"].startswith("depth-maps/sec fwd, 5-view 512") and j["config"]["workload"].startswith("cfg2")
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    quiet = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                           capture_output=True, text=True, timeout=60, env=dict(env, RANK="1", WORLD_SIZE="2"))
    assert quiet.returncode == 0 and not quiet.stdout.strip()


def test_fp16_arithmetic_falls_back_when_weights_leave_its_range():
    """The default '2xfp16' convolution arithmetic holds values below 65504; a model whose folded weights do not fit runs with
    three bf16 terms instead (engine._precision), with one warning."""
    import types
    import warnings
    from mvster_b200.engine import InferenceEngine
    assert packing.fp16_range_ok([torch.randn(3, 3)]) and packing.fp16_range_ok([])
    assert not packing.fp16_range_ok([torch.tensor([1.0, -4.0e4])]) and not packing.fp16_range_ok([torch.tensor([float("inf")])])
    eng = InferenceEngine(torch.device("cpu"))
    net = types.SimpleNamespace(fpn_precision="2xfp16", reg_precision="2xfp16")
    assert eng._precision(net, "fpn") == "2xfp16" and eng._precision(net, "reg") == "2xfp16"
    eng.fp16_safe["reg"] = False
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert eng._precision(net, "reg") == "3xbf16" and eng._precision(net, "reg") == "3xbf16" and eng._precision(net, "fpn") == "2xfp16"
    assert len(w) == 1
    net.reg_precision = "fp32"
    assert eng._precision(net, "reg") == "fp32"
    sd = build_model(SHIPPED, 0).state_dict()   # the benchmark's synthetic weights are far inside the range
    from mvster_b200 import fpn_engine
    fp = fpn_engine.pack_fpn({k: v for k, v in sd.items() if k.startswith("feature.")})
    assert packing.fp16_range_ok(v for k, v in fp.items() if k.endswith(".w"))


def test_python_flag_constants_match_the_header():
    import re
    from mvster_b200 import capi
    header = (REPO / "include" / "mvster_b200.h").read_text()
    for name in ("PARTIAL", "ACCUMULATE", "GENERIC", "NO_FUSE_D", "SQDIFF", "WINDOW", "NO_WINDOW"):
        assert getattr(capi, "ET_" + name) == int(re.search(rf"#define MVSTER_ET_{name} (\d+)", header).group(1)), name
    assert capi.TC3_FP16X2 == int(re.search(r"#define MVSTER_TC3_FP16X2 (\d+)", header).group(1))
    assert capi.MAX_VIEWS == int(re.search(r"#define MVSTER_MAX_VIEWS (\d+)", header).group(1))



def test_engine_input_contract_is_checked_before_any_launch():
    from mvster_b200 import engine, synth
    m = build_model(SHIPPED, seed=1)
    imgs, proj, dv = synth.make_inputs(2, 3, 64, 128, seed=0)
    engine.check_inputs(m, imgs, proj, dv)                                           # the contract itself passes
    engine.check_inputs(m, imgs, proj, torch.linspace(425, 935, 192).repeat(2, 1))   # the loaders' 192-entry list too
    bad = [
        (imgs[:1], proj, dv, "at least one source view"),
        ([i[:, :, :60] for i in imgs], proj, dv, "multiples of 64"),
        ([i[:, :, :, :96] for i in imgs], proj, dv, "multiples of 64"),
        (imgs[:2] + [imgs[2][:, :, :, :64]], proj, dv, "one shape"),
        ([i[:, :2] for i in imgs], proj, dv, "one shape"),
        (imgs, {k: v[:, :2] for k, v in proj.items()}, dv, "proj_matrices['stage1']"),
        (imgs, {k: v for k, v in proj.items() if k != "stage3"}, dv, "proj_matrices['stage3']"),
        (imgs, proj, dv[:1], "depth_values"),
        (imgs, proj, dv[:, :1], "depth_values"),
    ]
    for a, b, c, msg in bad:
        with pytest.raises(ValueError, match=msg.replace("[", r"\[").replace("]", r"\]")):
            engine.check_inputs(m, a, b, c)


def test_graph_output_copies_leave_no_cyclic_garbage():
    """engine._copy_outputs runs once per forward on ~45 MB of outputs: its copies must die with the caller's reference (no
    reference cycle that waits for the cycle collector) and keep the aliasing between the top-level and the stage-4 entries."""
    import gc
    import weakref
    from mvster_b200.engine import _copy_outputs
    gc.collect()
    gc.disable()
    try:
        t = torch.zeros(1000)
        st = {"depth": t, "mono_feat": torch.ones(3, 4).t()}
        out = {"stage4": st, "depth": t}
        r = _copy_outputs(out)
        assert r["depth"] is r["stage4"]["depth"] and r["depth"] is not t
        assert r["stage4"]["mono_feat"].stride() == st["mono_feat"].stride() and torch.equal(r["stage4"]["mono_feat"], st["mono_feat"])
        w = weakref.ref(r["depth"])
        del r
        assert w() is None
    finally:
        gc.enable()
