"""Stand-alone check of the tcgen05 convolutions against the exact-fp32 CUDA-core convolution.
Run one case per process (tests/test_gpu_tc_conv.py does that with a timeout) so that a device-side
trap or a barrier deadlock in a tensor-core kernel cannot take the rest of the suite down.

    python tests/tc_conv_check.py {v1|v2} CIN COUT KD B D H W NPASS [skip] [norelu]
    python tests/tc_conv_check.py {reg2d|reg2dv2} G B D H W NPASS
    python tests/tc_conv_check.py v3 CIN COUT KD K STRIDE B D H W [skip] [norelu] [h16 | b16]   (h16: two-fp16-term arithmetic;
    python tests/tc_conv_check.py d3 CIN COUT B D H W [skip] [h16 | b16]                         b16: one bf16 term + per-channel scale)
    ... b16 p16 / b16 p16f: packed bf16 operands (mvster_conv_tc3_pb16: x and skip octet-planar bf16), output packed bf16 / fp32
    ... h16 p16 / h16 p16f: packed fp16-pair operands (x and skip as (a1, a2) planes), output packed / fp32
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from mvster_b200 import capi, packing  # noqa: E402


def timeit(fn, n):
    for _ in range(3):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3  # us


def reg2d_main(gen):
    """whole reg2d U-Net, tensor-core 3x3x3 layers vs all-CUDA-core."""
    from util import SHIPPED, build_model
    G, B, D, H, W, npass = map(int, sys.argv[2:8])
    dev = torch.device("cuda", 0)
    sd = build_model(SHIPPED, 5).state_dict()
    packed = packing.pack_reg2d(sd, "reg.0" if G == 8 else "reg.3", capi.reg2d_layer_table(G))
    rng = np.random.RandomState(G + H)
    cost = torch.from_numpy((rng.randn(B, D, H, W, G) * 0.05).astype(np.float32)).to(dev)
    blob, tcb = packed["blob"].to(dev), packed["tc2_blob"].to(dev)
    want = capi.reg2d(blob, cost)
    got = capi.reg2d(blob, cost, tc_blob=tcb, npass=npass, kernel_gen=gen)
    torch.cuda.synchronize()
    err, scale = (got - want).abs().max().item(), want.abs().max().item()
    ws = torch.empty(capi.reg2d_workspace_floats(B, D, H, W), device=dev)
    out = torch.empty_like(want)
    t_simt = timeit(lambda: capi.reg2d(blob, cost, workspace=ws, out=out), 10)
    t_tc = timeit(lambda: capi.reg2d(blob, cost, workspace=ws, out=out, tc_blob=tcb, npass=npass, kernel_gen=gen), 10)
    print(json.dumps({"case": sys.argv[1:], "abs_err": err, "scale": scale, "rel": err / scale,
                      "finite": bool(torch.isfinite(got).all()), "us_tc": t_tc, "us_simt": t_simt}))


def bf16_round(t):
    return t.to(torch.bfloat16).to(t.dtype)


def v3_main():
    """generation 3 (persistent, 3 x bf16): v3 CIN COUT KD K STRIDE B D H W [skip] [norelu]; truth = fp64 torch conv."""
    import torch.nn.functional as F
    cin, cout, kd, k, stride, B, D, H, W = map(int, sys.argv[2:11])
    use_skip, relu = "skip" in sys.argv, "norelu" not in sys.argv
    dev = torch.device("cuda", 0)
    rng = np.random.RandomState(cin * 1000 + cout * 10 + kd + H + k + stride)
    x = torch.from_numpy(rng.randn(B, D, H, W, cin).astype(np.float32)).to(dev)
    w = torch.from_numpy((rng.randn(kd * k * k, cin, cout) / np.sqrt(kd * k * k * cin)).astype(np.float32))
    bias = torch.from_numpy(rng.randn(cout).astype(np.float32) * 0.1).to(dev)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    skip = torch.from_numpy(rng.randn(B, D, Ho, Wo, cout).astype(np.float32)).to(dev) if use_skip else None
    b16 = "b16" in sys.argv  # bf16-storage arithmetic: y = relu(scale * conv(bf16(x), bf16(w)) + bias) + skip, truth in fp64
    ch_scale = torch.from_numpy(rng.uniform(0.5, 2.0, cout).astype(np.float32)).to(dev) if b16 else None
    if b16:
        w = bf16_round(w)
    w5 = w.to(dev).double().reshape(kd, k, k, cin, cout).permute(4, 3, 0, 1, 2)
    xin = bf16_round(x) if b16 else x
    want = F.conv3d(xin.double().permute(0, 4, 1, 2, 3), w5, None, stride=(1, stride, stride), padding=(kd // 2, k // 2, k // 2))
    want = want.permute(0, 2, 3, 4, 1)
    if b16:
        want = want * ch_scale.double()
    want = want + bias.double()
    if relu:
        want = want.clamp_min(0)
    if skip is not None:
        want = want + skip.double()
    split = 1 if b16 else 2 if "h16" in sys.argv else 3  # h16: two fp16 terms per operand (MVSTER_TC3_FP16X2)
    wp = packing.pack_tc3_weights(w, kd, k, stride, split).to(dev)
    packed, packed_out = "p16" in sys.argv or "p16f" in sys.argv, "p16" in sys.argv
    extra = {}
    if packed:  # packed operands: the skip tensor is packed too (bf16: rounded; fp16 pair: 22 bits)
        assert b16 or split == 2
        pack, unpack = (capi.to_pb16, capi.from_pb16) if b16 else (capi.to_ph16, capi.from_ph16)
        if skip is not None and b16:
            want = want - skip.double() + bf16_round(skip).double()
        xp, sp = pack(x), None if skip is None else pack(skip)
        raw = lambda: capi.conv_tc3_pb16(xp, wp, bias, cout, kd, k, stride, relu, skip=sp, scale=ch_scale, out_pb16=packed_out)
        run = (lambda: unpack(raw())) if packed_out else raw
    else:
        run = lambda: capi.conv_tc3(x, wp, bias, cout, kd, k, stride, relu, skip=skip, split=split, scale=ch_scale)
    got = run()
    torch.cuda.synchronize()
    if packed_out and b16:  # the stored output is rounded to bf16: equal to the rounded truth except where fp32 noise flips a rounding
        extra["flip_frac"] = (got.double() != bf16_round(want.float()).double()).double().mean().item()
        want = bf16_round(want.float()).double()
    err, scale = (got.double() - want).abs().max().item(), want.abs().max().item()
    t_tc = timeit(raw if packed else run, 20)
    flops = 2.0 * B * D * Ho * Wo * kd * k * k * cin * cout
    print(json.dumps({"case": sys.argv[1:], "abs_err": err, "scale": scale, "rel": err / scale, "finite": bool(torch.isfinite(got).all()),
                      "us_tc": t_tc, "tflops_tc": flops / (t_tc * 1e-6) / 1e12, **extra}))


def d3_main():
    """generation-3 transposed conv (1,3,3)/stride (1,2,2): d3 CIN COUT B D H W [skip]; truth = fp64 conv_transpose3d."""
    import torch.nn.functional as F
    cin, cout, B, D, H, W = map(int, sys.argv[2:8])
    use_skip = "skip" in sys.argv
    dev = torch.device("cuda", 0)
    rng = np.random.RandomState(cin * 100 + cout + H)
    x = torch.from_numpy(rng.randn(B, D, H, W, cin).astype(np.float32)).to(dev)
    w = torch.from_numpy((rng.randn(9, cin, cout) / np.sqrt(2.25 * cin)).astype(np.float32))
    bias = torch.from_numpy(rng.randn(cout).astype(np.float32) * 0.1).to(dev)
    skip = torch.from_numpy(rng.randn(B, D, 2 * H, 2 * W, cout).astype(np.float32)).to(dev) if use_skip else None
    b16 = "b16" in sys.argv
    ch_scale = torch.from_numpy(rng.uniform(0.5, 2.0, cout).astype(np.float32)).to(dev) if b16 else None
    if b16:
        w = bf16_round(w)
    wt = w.to(dev).double().reshape(1, 3, 3, cin, cout).permute(3, 4, 0, 1, 2)  # [Cin][Cout][1][3][3]
    xin = bf16_round(x) if b16 else x
    want = F.conv_transpose3d(xin.double().permute(0, 4, 1, 2, 3), wt, None, stride=(1, 2, 2), padding=(0, 1, 1), output_padding=(0, 1, 1))
    want = want.permute(0, 2, 3, 4, 1)
    if b16:
        want = want * ch_scale.double()
    want = (want + bias.double()).clamp_min(0)
    if skip is not None:
        want = want + skip.double()
    split = 1 if b16 else 2 if "h16" in sys.argv else 3
    packed, packed_out = "p16" in sys.argv or "p16f" in sys.argv, "p16" in sys.argv
    extra = {}
    if packed:
        assert b16 or split == 2
        pack, unpack = (capi.to_pb16, capi.from_pb16) if b16 else (capi.to_ph16, capi.from_ph16)
        if skip is not None and b16:
            want = want - skip.double() + bf16_round(skip).double()
        xp, sp = pack(x), None if skip is None else pack(skip)
        pshape, pdt = capi._packed_out(B, D, cout, 2 * H, 2 * W, 1 if b16 else 2)
        shape, dt = (pshape, pdt) if packed_out else ((B, D, 2 * H, 2 * W, cout), torch.float32)
        buf = torch.full(shape, float("nan"), device=dev, dtype=dt)
        wps = [packing.pack_tc3_deconv_weights(w, r_, split).to(dev) for r_ in ((-1,) if 4 * cout <= 64 else (0, 1))]

        def raw():
            for r_, wp_ in zip((-1,) if 4 * cout <= 64 else (0, 1), wps):
                capi.deconv_tc3_pb16(xp, wp_, bias, cout, r_, True, skip=sp, scale=ch_scale, out_pb16=packed_out, out=buf)
            return buf
        run = (lambda: unpack(raw())) if packed_out else raw
    elif 4 * cout <= 64:
        wp = packing.pack_tc3_deconv_weights(w, -1, split).to(dev)
        run = lambda: capi.deconv_tc3(x, wp, bias, cout, -1, True, skip=skip, split=split, scale=ch_scale)
    else:
        wp0, wp1 = packing.pack_tc3_deconv_weights(w, 0, split).to(dev), packing.pack_tc3_deconv_weights(w, 1, split).to(dev)
        buf = torch.full((B, D, 2 * H, 2 * W, cout), float("nan"), device=dev)

        def run():
            capi.deconv_tc3(x, wp0, bias, cout, 0, True, skip=skip, out=buf, split=split, scale=ch_scale)
            return capi.deconv_tc3(x, wp1, bias, cout, 1, True, skip=skip, out=buf, split=split, scale=ch_scale)
    got = run()
    torch.cuda.synchronize()
    if packed_out and b16:
        extra["flip_frac"] = (got.double() != bf16_round(want.float()).double()).double().mean().item()
        want = bf16_round(want.float()).double()
    err, scale = (got.double() - want).abs().max().item(), want.abs().max().item()
    t_tc = timeit(raw if packed else run, 20)
    t_simt = timeit(lambda: capi.conv3d_ndhwc(x, w.to(dev), bias, 1, 1, 2, True, True, skip=skip), 20)
    print(json.dumps({"case": sys.argv[1:], "abs_err": err, "scale": scale, "rel": err / scale, "finite": bool(torch.isfinite(got).all()),
                      "us_tc": t_tc, "us_simt": t_simt, "tflops_tc": 2.0 * B * D * H * W * 9 * cin * cout / (t_tc * 1e-6) / 1e12, **extra}))


def main():
    mode = sys.argv[1]
    if mode == "v3":
        return v3_main()
    if mode == "d3":
        return d3_main()
    if mode.startswith("reg2d"):
        return reg2d_main(2)
    if mode != "v2":
        raise SystemExit(f"unknown mode {mode!r} (v2, v3, d3, reg2dv2)")
    cin, cout, kd, B, D, H, W, npass = map(int, sys.argv[2:10])
    use_skip, relu = "skip" in sys.argv, "norelu" not in sys.argv
    dev = torch.device("cuda", 0)
    rng = np.random.RandomState(cin * 1000 + cout * 10 + kd + H)
    x = torch.from_numpy(rng.randn(B, D, H, W, cin).astype(np.float32)).to(dev)
    w = torch.from_numpy((rng.randn(kd * 9, cin, cout) / np.sqrt(kd * 9 * cin)).astype(np.float32)).to(dev)
    bias = torch.from_numpy(rng.randn(cout).astype(np.float32) * 0.1).to(dev)
    skip = torch.from_numpy(rng.randn(B, D, H, W, cout).astype(np.float32)).to(dev) if use_skip else None
    want = capi.conv3d_ndhwc(x, w, bias, kd, 1, 1, False, relu, skip=skip)
    wp = packing.pack_tc2_weights(w.cpu(), npass).to(dev)
    run = lambda: capi.conv3d_tc2(x, wp, bias, cout, kd, relu, skip=skip, npass=npass)
    got = run()
    torch.cuda.synchronize()
    err, scale = (got - want).abs().max().item(), want.abs().max().item()
    # timing: back-to-back launches (activations stay in L2: a kernel-quality number, not a bench line)
    t_tc = timeit(run, 20)
    t_simt = timeit(lambda: capi.conv3d_ndhwc(x, w, bias, kd, 1, 1, False, relu, skip=skip), 20)
    flops = 2.0 * B * D * H * W * kd * 9 * cin * cout
    print(json.dumps({"case": sys.argv[1:], "abs_err": err, "scale": scale, "rel": err / scale, "finite": bool(torch.isfinite(got).all()),
                      "us_tc": t_tc, "us_simt": t_simt, "tflops_tc": flops / (t_tc * 1e-6) / 1e12}))


if __name__ == "__main__":
    main()
