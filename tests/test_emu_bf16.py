"""bf16-storage configuration (BASELINE configs[2]) on the CPU emulation: the SIMT kernels of the path (bf16 warp + ET kernels,
the bf16-input conv0, the cast, the head's bf16 `prob` input) run as compiled sources on host threads; the tcgen05 layers are
stood in for by tests/emu/emu_stubs.cpp, which DECODES THE PACKED WEIGHT SLABS with the library's own layer plan and evaluates
the layer directly - so the packers (operand terms, slab order, parity classes of the transposed layers), the blob offsets and
the layer chaining of mvster_reg2d_bf16 / mvster_reg2d_tc3_ex_f32 are checked here without a GPU.

Yardstick: oracle/mvster_oracle.py under ``storage(torch.bfloat16, fpn_internal=False)`` - "parity unpinned" (the reference has
no bf16 path; the definition is this repository's, see the oracle's header)."""
import numpy as np
import pytest
import torch

from test_emu_kernels import emu, emu_lib, from_ndhwc, nhwc  # noqa: F401  (fixtures)
from util import SHIPPED, build_model, narrow_et_inputs, oracle, oracle_cfg
from mvster_b200 import capi, packing, synth

BF16_ULP = 2.0 ** -7  # largest relative spacing of bf16 values (8 significand bits): a flipped rounding moves a value by at most this much


def q(t):
    return t.to(torch.bfloat16).float()


def test_cast_kernel_is_torchs_bf16_rounding(emu):
    rng = np.random.RandomState(0)
    x = torch.from_numpy(np.concatenate([rng.randn(1003).astype(np.float32) * 10 ** rng.uniform(-6, 6, 1003).astype(np.float32),
                                         np.array([0.0, -0.0, 1.0, 1.00390625, 1.01171875, 65504.0, 3.3e38, -3.3e38, np.inf, -np.inf], np.float32)]))
    got = capi.cast_bf16(x)
    assert got.dtype == torch.bfloat16 and torch.equal(got.view(torch.int16), x.to(torch.bfloat16).view(torch.int16))


ET_BF16 = [  # (B, nv, C, G, D, H, W, step_deg, rel_span, window)
    (1, 3, 64, 8, 8, 4, 8, 1.0, 0.5, None),     # stage-1 class: per-hypothesis tiled kernel
    (1, 3, 8, 4, 4, 8, 64, 1.0, 0.3, True),     # window kernel
    (1, 3, 16, 4, 4, 8, 32, 2.0, 0.2, True),
    (2, 2, 32, 8, 8, 4, 16, 4.0, 0.06, True),
    (1, 3, 8, 4, 4, 8, 32, 1.0, 0.3, False),    # tiled kernels for the small classes too
    (1, 2, 16, 4, 4, 8, 16, 2.0, 0.9, None),    # wide hypotheses: the window kernel falls back to its per-hypothesis gather
]


@pytest.mark.parametrize("case", ET_BF16)
def test_bf16_et_kernels_on_cpu_match_oracle(emu, case):
    B, nv, C_, G, D, H, W, step, span, window = case
    feats, cams, hypo = narrow_et_inputs(B, nv, C_, D, H, W, step, span, seed=3)
    feats = [q(f) for f in feats]                                        # what a bf16 pyramid output holds
    want = oracle.et_aggregate(feats, cams, hypo, True, G, 2.0)          # fp32 arithmetic on those values ...
    got = capi.et_fuse_bf16(nhwc(feats[0]).to(torch.bfloat16), [nhwc(f).to(torch.bfloat16) for f in feats[1:]], capi.pose(cams), hypo, G, 2.0,
                            window=window)
    assert "bf16" in capi.et_last_kernel(), capi.et_last_kernel()
    assert got.dtype == torch.bfloat16
    got = from_ndhwc(got.float())
    scale = want.abs().max().item()
    # ... rounded once on the way out: equal to the rounded oracle except where fp32 noise (2e-4 of max allowed, 2e-6 typical)
    # flips a rounding
    assert (got - q(want)).abs().max().item() <= BF16_ULP * scale + 2e-4 * scale
    assert (got == q(want)).float().mean().item() > 0.97


def test_bf16_et_interleaved_layout_on_cpu(emu):
    B, nv, C_, G, D, H, W = 1, 3, 8, 4, 4, 8, 32
    feats, cams, hypo = narrow_et_inputs(B, nv, C_, D, H, W, 1.0, 0.3, seed=5)
    feats = [q(f) for f in feats]
    perm = capi.interleave_perm(C_, G)
    a = capi.et_fuse_bf16(nhwc(feats[0]).to(torch.bfloat16), [nhwc(f).to(torch.bfloat16) for f in feats[1:]], capi.pose(cams), hypo, G, 2.0, window=True)
    il = [nhwc(f)[..., perm].contiguous().to(torch.bfloat16) for f in feats]
    b = capi.et_fuse_bf16(il[0], il[1:], capi.pose(cams), hypo, G, 2.0, window=True, interleaved=True)
    assert "interleaved" in capi.et_last_kernel()
    d = (a.float() - b.float()).abs().max().item()
    assert d <= BF16_ULP * a.float().abs().max().item()


def test_bf16_et_rejects_what_it_does_not_cover(emu):
    feats, cams, hypo = narrow_et_inputs(1, 2, 16, 4, 8, 8, 2.0, 0.3)
    bf = [nhwc(f).to(torch.bfloat16) for f in feats]
    with pytest.raises(Exception, match="unsupported"):
        capi.et_fuse_bf16(bf[0], bf[1:], capi.pose(cams), hypo, 8, 2.0)  # (C, G, D) = (16, 8, 4): no bf16 specialisation


def reg2d_case(k, D, H, W, seed):
    sd = build_model(SHIPPED, 5).state_dict()
    G = SHIPPED["group_cor_dim"][k]
    rng = np.random.RandomState(seed)
    cost = q(torch.from_numpy((rng.randn(1, G, D, H, W) * 0.1).astype(np.float32)))
    hypo = oracle.hypo_init_inverse(torch.tensor([[425.0, 935.0]]), D, H, W)
    return sd, G, cost, hypo


@pytest.mark.parametrize("k,D,H,W", [(0, 8, 8, 8), (3, 4, 16, 8), (2, 4, 24, 16)])
def test_bf16_reg2d_and_head_on_cpu_match_oracle(emu, k, D, H, W):
    """mvster_reg2d_bf16 (conv0 kernel as compiled source; conv1..conv11 through the decoded split-1 slab streams) + the head with
    its bf16 `prob` input against the oracle's reg2d under bf16 storage."""
    sd, G, cost, hypo = reg2d_case(k, D, H, W, k)
    with torch.no_grad(), oracle.storage(torch.bfloat16, False):
        logits = oracle.reg2d_logits(sd, f"reg.{k}", cost)
        want = oracle.depth_head(logits, hypo, k, 0.5, True)
    with torch.no_grad():
        exact = oracle.depth_head(oracle.reg2d_logits(sd, f"reg.{k}", cost), hypo, k, 0.5, True)  # the fp32 network on the same input
    packed = packing.pack_reg2d_bf16(sd, f"reg.{k}", capi.reg2d_layer_table(G))
    feat8 = capi.reg2d_bf16(packed["blob_q"], packed["tc3_blob"], packed["scales"], cost.permute(0, 2, 3, 4, 1).contiguous().to(torch.bfloat16))
    h = capi.head(hypo, 0.5, feat8=feat8, prob_w=packed["prob_w"], prob_b=packed["prob_b"], bf16_input=True)
    err = (h["attn_weight"] - want["attn_weight"]).abs().max().item()
    floor = (exact["attn_weight"] - want["attn_weight"]).abs().max().item()  # how far bf16 storage itself moves the probabilities
    # a handful of activations round the other way (fp32 accumulation order): the distance to the bf16 oracle must stay well below
    # the distance between the bf16 and the fp32 network
    assert err <= 0.25 * floor + 1e-4, (err, floor)
    gap = want["attn_weight"].topk(2, dim=1).values
    stable = (gap[:, 0] - gap[:, 1]) > 4 * err + 1e-3
    assert torch.equal(h["depth"][stable], want["depth"][stable])


def test_packed_and_container_bf16_reg2d_agree_on_cpu(emu):
    """The two drivers of mvster_reg2d_bf16 (packed bf16 operands between the layers / bf16-rounded values in fp32 containers)
    chain the same layers with the same roundings: identical results (the tensor-core layers being the decoded-slab stand-ins)."""
    sd, G, cost, hypo = reg2d_case(3, 4, 16, 24, 21)
    packed = packing.pack_reg2d_bf16(sd, "reg.3", capi.reg2d_layer_table(G))
    cb = cost.permute(0, 2, 3, 4, 1).contiguous().to(torch.bfloat16)
    a = capi.reg2d_bf16(packed["blob_q"], packed["tc3_blob"], packed["scales"], cb, packed=True)
    b = capi.reg2d_bf16(packed["blob_q"], packed["tc3_blob"], packed["scales"], cb, packed=False)
    assert torch.isfinite(a).all() and torch.equal(a, b)


def test_packed_fp16_pair_reg2d_on_cpu(emu):
    """mvster_reg2d_tc3_ex_f32 with MVSTER_REG2D_PACKED (conv0 kernel writing the fp16 pair as compiled source, every later layer
    through the decoded slab streams on packed (a1, a2) planes) against the unpacked form and the oracle."""
    k, D, H, W = 3, 4, 16, 24
    sd, G, cost, hypo = reg2d_case(k, D, H, W, 31)
    cost = cost * 3 + torch.randn_like(cost) * 0.01  # not bf16-valued
    with torch.no_grad():
        want = oracle.depth_head(oracle.reg2d_logits(sd, f"reg.{k}", cost), hypo, k, 0.5, True)
    packed = packing.pack_reg2d(sd, f"reg.{k}", capi.reg2d_layer_table(G))
    cl = cost.permute(0, 2, 3, 4, 1).contiguous()
    a = capi.reg2d(packed["blob"], cl, tc_blob=packed["tc3h_blob"], kernel_gen=3, split=2, packed=True)
    b = capi.reg2d(packed["blob"], cl, tc_blob=packed["tc3h_blob"], kernel_gen=3, split=2, packed=False)
    assert torch.isfinite(a).all() and (a - b).abs().max().item() <= 2e-6 * b.abs().max().item()
    h = capi.head(hypo, 0.5, feat8=a, prob_w=packed["prob_w"], prob_b=packed["prob_b"])
    assert (h["attn_weight"] - want["attn_weight"]).abs().max().item() < 2e-5


def test_packed_conv0_reports_values_outside_the_fp16_terms_on_cpu(emu):
    """The range check of the two-fp16-term arithmetic moves with the split: in the packed form conv0 (and, on the GPU, every
    epilogue) checks what it STORES.  Here: the conv0 kernel as compiled source, flag word registered like the engine does."""
    import ctypes as C
    sd, G, cost, hypo = reg2d_case(3, 4, 8, 16, 41)
    packed = packing.pack_reg2d(sd, "reg.3", capi.reg2d_layer_table(G))
    flag = torch.zeros(1, dtype=torch.int32)
    lib = emu
    for scale, want in ((1.0, 0), (1.0e7, 1)):
        flag.zero_()
        lib.mvster_tc3_set_overflow_flag(C.c_void_p(flag.data_ptr()))
        try:
            capi.reg2d(packed["blob"], (cost * scale).permute(0, 2, 3, 4, 1).contiguous(), tc_blob=packed["tc3h_blob"], kernel_gen=3, split=2, packed=True)
        finally:
            lib.mvster_tc3_set_overflow_flag(None)
        assert int(flag.item()) == want, (scale, int(flag.item()))


def test_pb16_layout_helpers_round_trip():
    x = torch.randn(2, 3, 5, 7, 16)
    p = capi.to_pb16(x)
    assert p.shape == (2, 3, 2, 5, 7, 8) and p.dtype == torch.bfloat16
    assert torch.equal(capi.from_pb16(p), x.to(torch.bfloat16).float())
    assert torch.equal(p[1, 2, 1, 4, 6], x[1, 2, 4, 6, 8:].to(torch.bfloat16))
    h = capi.to_ph16(x * 50)
    assert h.shape == (2, 3, 2, 2, 5, 7, 8) and h.dtype == torch.float16
    assert torch.equal(h[1, 2, 1, 0, 4, 6], (x * 50)[1, 2, 4, 6, 8:].half())
    assert (capi.from_ph16(h) - x * 50).abs().max().item() <= 2.0 ** -21 * (x * 50).abs().max().item()


@pytest.mark.parametrize("split", [2, 3])
def test_reg2d_tc3_slab_streams_decode_to_the_network_on_cpu(emu, split):
    """The fp32-faithful tensor-core reg2d (mvster_reg2d_tc3_ex_f32) with its layers evaluated from the DECODED slab streams
    ('tc3h_blob' two fp16 terms / 'tc3_blob' three bf16 terms): packer and driver against the oracle, on the CPU."""
    k, D, H, W = 3, 4, 16, 16
    sd, G, cost, hypo = reg2d_case(k, D, H, W, 11)
    with torch.no_grad():
        want = oracle.depth_head(oracle.reg2d_logits(sd, f"reg.{k}", cost), hypo, k, 0.5, True)
    packed = packing.pack_reg2d(sd, f"reg.{k}", capi.reg2d_layer_table(G))
    feat8 = capi.reg2d(packed["blob"], cost.permute(0, 2, 3, 4, 1).contiguous(), tc_blob=packed["tc3h_blob" if split == 2 else "tc3_blob"],
                       kernel_gen=3, split=split)
    h = capi.head(hypo, 0.5, feat8=feat8, prob_w=packed["prob_w"], prob_b=packed["prob_b"])
    assert (h["attn_weight"] - want["attn_weight"]).abs().max().item() < 2e-5


def test_bf16_storage_engine_forward_on_cpu(emu, monkeypatch):
    """MVS4net.forward with storage='bf16' end to end on the emulation (native fp32 pyramid -> cast -> bf16 warp/ET -> bf16
    regulariser -> head) against oracle.cascade_forward(storage_dtype=bfloat16, storage_fpn_internal=False)."""
    import contextlib
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)
    model = build_model(SHIPPED, 2)
    model.fpn_backend, model.fpn_precision, model.use_cuda_graph, model.overlap_stages = "native", "fp32", False, False
    model.storage = "bf16"
    imgs, proj, dv = synth.make_inputs(1, 3, 64, 128, seed=4)
    from mvster_b200.engine import InferenceEngine
    eng = InferenceEngine(torch.device("cpu"))
    eng.refresh_weights(model)
    with torch.no_grad():
        out = eng.forward(model, imgs, proj, dv)
        want = oracle.cascade_forward(model.state_dict(), oracle_cfg(SHIPPED), imgs, proj, dv, storage_dtype=torch.bfloat16,
                                      storage_fpn_internal=False)
        exact = oracle.cascade_forward(model.state_dict(), oracle_cfg(SHIPPED), imgs, proj, dv)
    # End to end the comparison is statistical by nature: 0.15 % of the pyramid's outputs round the other way (fp32 noise of the
    # pyramid), which moves most cost values a little, and every later bf16 rounding of the regulariser then decorrelates - the
    # per-kernel tests above are the sharp ones.  Required here: our probabilities are no farther from the bf16 oracle than the fp32
    # network is (mean over all voxels, every stage) and the depth agrees on the pixels whose top-2 gap exceeds that noise.
    for s_ in range(1, 5):
        a, b, e = out[f"stage{s_}"]["attn_weight"], want[f"stage{s_}"]["attn_weight"], exact[f"stage{s_}"]["attn_weight"]
        assert (a - b).abs().mean().item() <= 1.25 * (e - b).abs().mean().item(), s_  # measured 0.5 .. 0.87
    from oracle.compare import cascade_parity
    rep = cascade_parity(out, want, tie_gap=0.1, max_bad=5e-2, max_attn1=1.0)
    assert rep["ok"], rep
    assert out["stage4"]["mono_feat"].shape == want["stage4"]["mono_feat"].shape
    assert torch.equal(out["stage4"]["mono_feat"], want["stage4"]["mono_feat"]) or \
        (out["stage4"]["mono_feat"] - want["stage4"]["mono_feat"]).abs().max().item() <= BF16_ULP * want["stage4"]["mono_feat"].abs().max().item()
