"""bf16-storage configuration (BASELINE configs[2]) on the GPU, through the C ABI: one-term tcgen05 convolutions against an fp64
convolution of the bf16-rounded operands, the bf16 warp + ET kernels, the bf16 regulariser + head and the whole forward with
``model.storage = "bf16"`` against ``oracle.cascade_forward(storage_dtype=torch.bfloat16, storage_fpn_internal=False)``.

PARITY UNPINNED for this configuration (the reference has no bf16 path: its pixel grid is hard-coded fp32,
mvs4net_utils.py:28-29); the definition is this repository's, see oracle/mvster_oracle.py `storage`.  Runs last in the suite."""
import json
import subprocess
import sys

import numpy as np
import pytest
import torch

from util import REPO, SHIPPED, build_model, narrow_et_inputs, oracle, oracle_cfg
from mvster_b200 import capi, packing, synth

pytestmark = pytest.mark.gpu
BF16_ULP = 2.0 ** -7  # largest relative spacing of bf16 values (8 significand bits): a flipped rounding moves a value by at most this much


def q(t):
    return t.to(torch.bfloat16).float()


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


TC_CASES = [  # tests/tc_conv_check.py cases with b16: y = relu(scale * conv(bf16(x), bf16(w)) + bias) [+ skip], fp64 yardstick
    "v3 16 16 3 3 1 1 8 24 40 b16",           # depth taps, ragged tiles
    "v3 32 32 3 3 1 2 4 16 16 norelu b16",    # two channel chunks, batch 2
    "v3 64 64 3 3 1 1 4 64 80 skip b16",      # conv6 at cfg2 stage 4
    "v3 8 16 1 3 2 1 4 64 80 b16",            # conv1: stride (1,2,2), two taps per MMA
    "v3 16 32 1 3 2 1 4 30 44 b16",           # conv3, odd output size
    "v3 16 16 3 3 1 1 4 256 320 b16",         # conv2 at cfg2 stage 4
    "d3 16 8 1 2 24 40 skip b16",             # conv11 class
    "d3 32 16 2 2 16 16 b16",                 # conv9 class
    "d3 64 32 1 4 8 10 skip b16",             # conv7 class: two launches
    # packed operands (mvster_conv_tc3_pb16 / mvster_deconv_tc3_pb16: x and skip octet-planar bf16, TMA straight into the operand
    # ring).  p16f = fp32 output: the sharp check of the load / MMA path; p16 = packed bf16 output (rounded once)
    "v3 16 16 3 3 1 1 8 24 40 skip b16 p16f",  # depth taps, ragged tiles, 4-D merged map
    "v3 8 16 1 3 2 1 4 64 80 b16 p16f",        # conv1: stride 2 (5-D map with element strides), one octet, two taps per MMA
    "v3 16 32 1 3 2 1 4 30 44 b16 p16f",       # conv3, odd output size
    "v3 32 64 1 3 2 2 2 32 32 b16 p16f",       # conv5: two stages per parity class
    "v3 64 64 3 3 1 1 4 64 80 b16 p16f",       # conv6
    "v3 32 32 3 3 1 2 4 16 16 norelu b16 p16", # packed output, batch 2
    "v3 16 16 3 3 1 1 4 256 320 b16 p16",      # conv2 at cfg2 stage 4
    "v3 8 16 1 3 2 1 4 512 640 b16 p16",       # conv1 at cfg2 stage 4
    "d3 16 8 1 2 24 40 skip b16 p16f",         # conv11: packed skip (8 channels = NHWC bf16), fp32 out
    "d3 32 16 2 2 16 16 skip b16 p16",         # conv9
    "d3 64 32 1 4 8 10 skip b16 p16",          # conv7: two launches into one packed output
]


@pytest.mark.parametrize("case", TC_CASES)
def test_one_term_bf16_tc_conv_matches_fp64_conv_of_rounded_operands(case):
    p = subprocess.run([sys.executable, str(REPO / "tests" / "tc_conv_check.py"), *case.split()], capture_output=True, text=True, timeout=90)
    assert p.returncode == 0, f"subprocess failed:\n{p.stdout[-2000:]}\n{p.stderr[-3000:]}"
    res = json.loads(p.stdout.strip().splitlines()[-1])
    with open(REPO / "gpurun_out" / "tc_conv_report.jsonl", "a") as f:
        f.write(json.dumps(res) + "\n")
    if case.split()[-1] == "p16":  # packed bf16 output: the rounded truth, except where fp32 summation noise flips a rounding
        assert res["finite"] and res["rel"] <= BF16_ULP and res["flip_frac"] < 2e-3, res
        return
    # products of bf16 values are exact in fp32: what remains is the tensor core's fp32 accumulation of K <= 1728 terms
    assert res["finite"] and res["rel"] < 1e-5, res


def test_cast_kernel_is_torchs_bf16_rounding():
    rng = np.random.RandomState(0)
    x = torch.from_numpy((rng.randn(1 << 20) * 10 ** rng.uniform(-6, 6, 1 << 20)).astype(np.float32)).cuda()
    x[:6] = torch.tensor([0.0, 1.00390625, 1.01171875, 65504.0, float("inf"), -float("inf")])
    for n in (1 << 20, 1003, 8, 7):
        got = capi.cast_bf16(x[:n].clone())
        assert torch.equal(got.view(torch.int16), x[:n].to(torch.bfloat16).view(torch.int16)), n


ET_BF16 = [  # (B, nv, C, G, D, H, W, step_deg, rel_span, window)
    (1, 3, 64, 8, 8, 16, 24, 1.0, 0.5, None),
    (1, 3, 8, 4, 4, 37, 70, 1.0, 0.3, True),
    (1, 5, 16, 4, 4, 17, 48, 2.0, 0.2, True),
    (2, 4, 32, 8, 8, 7, 33, 4.0, 0.06, True),
    (1, 3, 8, 4, 4, 16, 40, 1.0, 0.3, False),
    (1, 2, 16, 4, 4, 8, 16, 2.0, 0.9, None),
]


@pytest.mark.parametrize("case", ET_BF16)
def test_bf16_et_kernels_match_oracle(case):
    B, nv, C_, G, D, H, W, step, span, window = case
    feats, cams, hypo = narrow_et_inputs(B, nv, C_, D, H, W, step, span, seed=3)
    feats = [q(f) for f in feats]
    want = oracle.et_aggregate(feats, cams, hypo, True, G, 2.0)
    got = capi.et_fuse_bf16(nhwc(feats[0]).cuda().to(torch.bfloat16), [nhwc(f).cuda().to(torch.bfloat16) for f in feats[1:]],
                            capi.pose(cams.cuda()), hypo.cuda(), G, 2.0, window=window)
    assert "bf16" in capi.et_last_kernel(), capi.et_last_kernel()
    got = got.float().permute(0, 4, 1, 2, 3).cpu()
    scale = want.abs().max().item()
    assert torch.isfinite(got).all()
    assert (got - q(want)).abs().max().item() <= BF16_ULP * scale + 2e-4 * scale
    assert (got == q(want)).float().mean().item() > 0.97
    # interleaved channel layout: same values
    if C_ // G in (2, 4) and window:
        perm = capi.interleave_perm(C_, G)
        il = [nhwc(f)[..., perm].contiguous().cuda().to(torch.bfloat16) for f in feats]
        b = capi.et_fuse_bf16(il[0], il[1:], capi.pose(cams.cuda()), hypo.cuda(), G, 2.0, window=True, interleaved=True)
        assert "interleaved" in capi.et_last_kernel()
        assert (b.float().permute(0, 4, 1, 2, 3).cpu() - got).abs().max().item() <= BF16_ULP * scale


@pytest.mark.parametrize("k,D,H,W", [(0, 8, 8, 8), (3, 4, 16, 8), (2, 4, 24, 16), (3, 4, 64, 80), (1, 8, 40, 56)])
def test_packed_and_container_bf16_reg2d_are_the_same_arithmetic(k, D, H, W):
    """mvster_reg2d_bf16 with packed bf16 operands between the layers (no conversion pass) against the same network with the
    bf16-rounded activations kept in fp32 containers: same MMAs, same epilogue arithmetic, same roundings - bit-identical."""
    sd = build_model(SHIPPED, 5).state_dict()
    G = SHIPPED["group_cor_dim"][k]
    rng = np.random.RandomState(100 + k)
    cost = torch.from_numpy((rng.randn(2, D, H, W, G) * 0.1).astype(np.float32)).cuda().to(torch.bfloat16)
    packed = {n: t.cuda() for n, t in packing.pack_reg2d_bf16(sd, f"reg.{k}", capi.reg2d_layer_table(G)).items()}
    a = capi.reg2d_bf16(packed["blob_q"], packed["tc3_blob"], packed["scales"], cost, packed=True)
    b = capi.reg2d_bf16(packed["blob_q"], packed["tc3_blob"], packed["scales"], cost, packed=False)
    torch.cuda.synchronize()
    assert torch.isfinite(a).all() and torch.equal(a, b), (a - b).abs().max().item()


@pytest.mark.parametrize("k,D,H,W", [(0, 8, 8, 8), (3, 4, 16, 8), (2, 4, 24, 16), (3, 4, 64, 80), (1, 8, 40, 56), (3, 4, 256, 320)])
def test_packed_fp16_pair_reg2d_equals_the_unpacked_form(k, D, H, W):
    """The fp32-faithful default regulariser (two fp16 terms per operand) with PACKED activations between the layers - the producing
    epilogue stores the next layer's operand terms (a1, a2) - against the form whose consumers split fp32 activations themselves:
    same products; only the skip sums see their addend to 22 instead of 24 bits."""
    sd = build_model(SHIPPED, 5).state_dict()
    G = SHIPPED["group_cor_dim"][k]
    rng = np.random.RandomState(200 + k)
    cost = torch.from_numpy((rng.randn(1, D, H, W, G) * 0.3).astype(np.float32)).cuda()
    packed = {n: t.cuda() for n, t in packing.pack_reg2d(sd, f"reg.{k}", capi.reg2d_layer_table(G)).items()}
    a = capi.reg2d(packed["blob"], cost, tc_blob=packed["tc3h_blob"], kernel_gen=3, split=2, packed=True)
    b = capi.reg2d(packed["blob"], cost, tc_blob=packed["tc3h_blob"], kernel_gen=3, split=2, packed=False)
    exact = capi.reg2d(packed["blob"], cost)  # every layer in exact fp32 on the CUDA cores
    torch.cuda.synchronize()
    scale = exact.abs().max().item()
    assert torch.isfinite(a).all()
    assert (a - b).abs().max().item() <= 2e-6 * scale, (a - b).abs().max().item() / scale
    assert (a - exact).abs().max().item() <= 2e-5 * scale, (a - exact).abs().max().item() / scale


def test_packed_fp16_pair_layers_report_values_outside_the_fp16_terms():
    """The range check of the two-fp16-term arithmetic (mvster_tc3_set_overflow_flag) in the packed form: the producer of a packed
    tensor checks what it stores - conv0 (CUDA cores) and every tensor-core epilogue."""
    from mvster_b200 import _lib
    lib = _lib.load()
    sd = build_model(SHIPPED, 5).state_dict()
    packed = {n: t.cuda() for n, t in packing.pack_reg2d(sd, "reg.3", capi.reg2d_layer_table(4)).items()}
    rng = np.random.RandomState(7)
    cost = torch.from_numpy((rng.randn(1, 4, 32, 48, 4) * 0.3).astype(np.float32)).cuda()
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")

    def run(fn):
        flag.zero_()
        lib.mvster_tc3_set_overflow_flag(flag.data_ptr())
        try:
            fn()
            torch.cuda.synchronize()
        finally:
            lib.mvster_tc3_set_overflow_flag(None)
        return int(flag.item())

    reg = lambda c: capi.reg2d(packed["blob"], c, tc_blob=packed["tc3h_blob"], kernel_gen=3, split=2, packed=True)
    assert run(lambda: reg(cost)) == 0
    assert run(lambda: reg(cost * 1.0e7)) == 1            # conv0's output leaves the range
    # one tensor-core layer alone: in-range input, weights that push the stored output beyond 65504
    x = capi.to_ph16(torch.full((1, 2, 16, 16, 16), 100.0, device="cuda"))
    w_big = packing.pack_tc3_weights(torch.full((9, 16, 16), 10.0), 1, 3, 1, 2).cuda()   # 144 * 100 * 10 = 144 000
    w_ok = packing.pack_tc3_weights(torch.full((9, 16, 16), 0.01), 1, 3, 1, 2).cuda()
    assert run(lambda: capi.conv_tc3_pb16(x, w_ok, None, 16, 1, 3, 1, True)) == 0
    assert run(lambda: capi.conv_tc3_pb16(x, w_big, None, 16, 1, 3, 1, True)) == 1
    assert run(lambda: capi.conv_tc3_pb16(x, w_big, None, 16, 1, 3, 1, True, out_pb16=False)) == 0  # fp32 output: nothing is split


@pytest.mark.parametrize("k,D,H,W", [(0, 8, 8, 8), (3, 4, 16, 8), (2, 4, 24, 16), (3, 4, 64, 80)])
def test_bf16_reg2d_and_head_match_oracle(k, D, H, W):
    sd = build_model(SHIPPED, 5).state_dict()
    G = SHIPPED["group_cor_dim"][k]
    rng = np.random.RandomState(k)
    cost = q(torch.from_numpy((rng.randn(1, G, D, H, W) * 0.1).astype(np.float32)))
    hypo = oracle.hypo_init_inverse(torch.tensor([[425.0, 935.0]]), D, H, W)
    with torch.no_grad(), oracle.storage(torch.bfloat16, False):
        want = oracle.depth_head(oracle.reg2d_logits(sd, f"reg.{k}", cost), hypo, k, 0.5, True)
    with torch.no_grad():
        exact = oracle.depth_head(oracle.reg2d_logits(sd, f"reg.{k}", cost), hypo, k, 0.5, True)
    packed = {n: t.cuda() for n, t in packing.pack_reg2d_bf16(sd, f"reg.{k}", capi.reg2d_layer_table(G)).items()}
    feat8 = capi.reg2d_bf16(packed["blob_q"], packed["tc3_blob"], packed["scales"], cost.permute(0, 2, 3, 4, 1).contiguous().cuda().to(torch.bfloat16))
    h = capi.head(hypo.cuda(), 0.5, feat8=feat8, prob_w=packed["prob_w"], prob_b=packed["prob_b"], bf16_input=True)
    attn = h["attn_weight"].cpu()
    err = (attn - want["attn_weight"]).abs().max().item()
    floor = (exact["attn_weight"] - want["attn_weight"]).abs().max().item()
    assert torch.isfinite(attn).all()
    d, f = (attn - want["attn_weight"]).abs(), (exact["attn_weight"] - want["attn_weight"]).abs()
    if H * W <= 24 * 16:
        # small volumes (~1e5 activations): no activation rounds the other way, or a handful do - far below what bf16 storage moves
        assert err <= 0.3 * floor + 1e-4, (err, floor)
        assert d.mean().item() <= 0.1 * f.mean().item() + 1e-6
    else:
        # Millions of activations: fp32 summation order (tensor core vs the oracle's CPU convolution, ~1e-7 relative) flips the bf16
        # rounding of ~1 activation in 4e4, and every flip re-draws the roundings downstream of it.  The CPU emulation of this test
        # case (fp64 accumulation, tests/emu) lands on the same numbers as the tensor core (max 0.0085, mean 2.9e-4 at 64x80), i.e.
        # the difference is the oracle's own summation noise.  Required: closer to the bf16 oracle than the fp32 network is.
        assert d.mean().item() <= 0.8 * f.mean().item(), (d.mean().item(), f.mean().item())
        assert err <= 1.5 * floor, (err, floor)
    gap = want["attn_weight"].topk(2, dim=1).values
    stable = (gap[:, 0] - gap[:, 1]) > 4 * err + 1e-3
    assert torch.equal(h["depth"].cpu()[stable], want["depth"][stable])


@pytest.mark.parametrize("graph,overlap", [(False, False), (True, True)])
def test_bf16_storage_forward_matches_bf16_oracle(graph, overlap):
    """MVS4net.forward with storage='bf16' (eager single-stream, and the default CUDA-graph two-stream schedule) against the bf16
    oracle.  Statistical by nature end to end (see tests/test_emu_bf16.py): our probabilities must be about as close to the bf16 oracle as
    the fp32 network is, or closer (a defect would put them an order of magnitude farther), and the depth must agree on the pixels whose top-2 probability gap exceeds that noise."""
    from oracle.compare import cascade_parity
    dev = torch.device("cuda", 0)
    model = build_model(SHIPPED, 2)
    sd = {k_: v.clone() for k_, v in model.state_dict().items()}
    model = model.to(dev)
    model.storage, model.use_cuda_graph, model.overlap_stages = "bf16", graph, overlap
    imgs, proj, dv = synth.make_inputs(1, 3, 128, 192, seed=4)
    with torch.no_grad():
        for _ in range(2 if graph else 1):  # second call = graph replay
            out = model([t.to(dev) for t in imgs], {k_: v.to(dev) for k_, v in proj.items()}, dv.to(dev))
        torch.cuda.synchronize()
        want = oracle.cascade_forward(sd, oracle_cfg(SHIPPED), imgs, proj, dv, storage_dtype=torch.bfloat16, storage_fpn_internal=False)
        exact = oracle.cascade_forward(sd, oracle_cfg(SHIPPED), imgs, proj, dv)
    for s_ in range(1, 5):
        a, b, e = out[f"stage{s_}"]["attn_weight"].cpu(), want[f"stage{s_}"]["attn_weight"], exact[f"stage{s_}"]["attn_weight"]
        assert torch.isfinite(a).all()
        assert (a - b).abs().mean().item() <= 1.25 * (e - b).abs().mean().item(), s_  # measured 0.5 .. 0.87
        assert out[f"stage{s_}"]["depth"].dtype == torch.float32
    rep = cascade_parity(out, want, tie_gap=0.1, max_bad=5e-2, max_attn1=1.0)
    assert rep["ok"], rep
    mono, wm = out["stage4"]["mono_feat"].cpu(), want["stage4"]["mono_feat"]
    assert mono.shape == wm.shape and (mono - wm).abs().max().item() <= BF16_ULP * wm.abs().max().item()
    assert (mono == wm).float().mean().item() > 0.99


def test_bf16_storage_rejects_what_it_does_not_cover():
    dev = torch.device("cuda", 0)
    cfg = dict(SHIPPED, reg_net="reg3d")
    model = build_model(cfg, 3).to(dev)
    model.storage, model.use_cuda_graph = "bf16", False
    imgs, proj, dv = synth.make_inputs(1, 2, 64, 64, seed=1)
    with torch.no_grad(), pytest.raises(NotImplementedError, match="bf16"):
        model([t.to(dev) for t in imgs], {k_: v.to(dev) for k_, v in proj.items()}, dv.to(dev))
