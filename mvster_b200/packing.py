"""Host-side weight preparation for the C ABI: fold eval-mode BatchNorm into the preceding
(transposed) convolution and lay the kernels out as ``[tap][Cin][Cout]`` fp32.

conv -> BN(eval):  y = (conv(x, w) - mean) / sqrt(var + eps) * gamma + beta
                     = conv(x, w * s) + (beta - mean * s),   s = gamma / sqrt(var + eps)
(mvs4net_utils.py:116-123, :885-898; eps = 1e-5, the nn.BatchNorm3d default).
"""
from __future__ import annotations

from typing import Dict, Mapping, Tuple

import torch

Tensor = torch.Tensor
BN_EPS = 1e-5


def bn_scale_shift(sd: Mapping[str, Tensor], p: str) -> Tuple[Tensor, Tensor]:
    s = sd[p + ".weight"].double() / torch.sqrt(sd[p + ".running_var"].double() + BN_EPS)
    return s, sd[p + ".bias"].double() - sd[p + ".running_mean"].double() * s


def fold_conv3d(w: Tensor, scale: Tensor, shift: Tensor) -> Tuple[Tensor, Tensor]:
    """Conv3d weight [Cout,Cin,kd,kh,kw] -> ([kd*kh*kw, Cin, Cout], [Cout]) fp32."""
    wf = w.double() * scale.view(-1, 1, 1, 1, 1)
    taps = w.shape[2] * w.shape[3] * w.shape[4]
    return wf.permute(2, 3, 4, 1, 0).reshape(taps, w.shape[1], w.shape[0]).float().contiguous(), shift.float().contiguous()


def fold_deconv3d(w: Tensor, scale: Tensor, shift: Tensor) -> Tuple[Tensor, Tensor]:
    """ConvTranspose3d weight [Cin,Cout,kd,kh,kw] -> ([kd*kh*kw, Cin, Cout], [Cout]) fp32."""
    wf = w.double() * scale.view(1, -1, 1, 1, 1)
    taps = w.shape[2] * w.shape[3] * w.shape[4]
    return wf.permute(2, 3, 4, 0, 1).reshape(taps, w.shape[0], w.shape[1]).float().contiguous(), shift.float().contiguous()


TF32_MASK = -8192  # 0xFFFFE000 as int32: keep sign, exponent and the top 10 mantissa bits


def pack_tc2_weights(w: Tensor, npass: int = 3) -> Tensor:
    """[taps][Cin][Cout] fp32 -> slabs for the staged-tile tcgen05 kernel (conv_tc2.cu):
    [hi | lo] x [taps][Cin/16][N][16] with N = max(Cout, 16) (zero rows pad Cout = 8 to the minimum UMMA N)."""
    taps, cin, cout = w.shape
    if cin % 16:
        raise ValueError(f"Cin={cin} is not a multiple of 16")
    n = 16 if cout < 16 else (80 if cout > 64 else cout)  # the UMMA N the kernel instantiates (16 / 32 / 64 / 80)
    wp = torch.zeros((taps, cin, n), dtype=torch.float32)
    wp[:, :, :cout] = w.float()
    slabs = wp.reshape(taps, cin // 16, 16, n).permute(0, 1, 3, 2).contiguous()
    if npass == 1:
        return slabs.reshape(-1)
    hi = (slabs.view(torch.int32) & TF32_MASK).view(torch.float32)
    lo = ((slabs - hi).view(torch.int32) & TF32_MASK).view(torch.float32)
    return torch.cat([hi.reshape(-1), lo.reshape(-1)])


def bf16_split3(w: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """fp32 -> three bf16 terms with w == t1 + t2 + t3 to 24 bits (round-to-nearest at every step, exact residuals)."""
    t1 = w.to(torch.bfloat16)
    r = w - t1.float()
    t2 = r.to(torch.bfloat16)
    t3 = (r - t2.float()).to(torch.bfloat16)
    return t1, t2, t3


def fp16_split2(w: Tensor) -> Tuple[Tensor, Tensor]:
    """fp32 -> two fp16 terms with w == t1 + 2^-11 t2 to 22 bits: t1 = fp16(w), t2 = fp16(2^11 (w - t1)) (the scaling keeps the
    residual out of fp16's subnormal range).  |w| must stay below 65504."""
    wc = w.clamp(-65504.0, 65504.0)
    t1 = wc.to(torch.float16)
    t2 = ((wc - t1.float()) * 2048.0).clamp(-65504.0, 65504.0).to(torch.float16)
    return t1, t2


FP16_WEIGHT_LIMIT = 3.0e4  # |w| above this leaves no headroom below fp16's 65504: such a model runs with three bf16 terms instead


def fp16_range_ok(tensors) -> bool:
    """True if every (finite) value of the given folded weight tensors stays inside the range the two-fp16-term arithmetic
    represents faithfully (``MVSTER_TC3_FP16X2``); the engine falls back to three bf16 terms otherwise."""
    for t in tensors:
        t = t.detach().float()
        if t.numel() and (not bool(torch.isfinite(t).all()) or float(t.abs().max()) > FP16_WEIGHT_LIMIT):
            return False
    return True


def _split_terms(wt: Tensor, split: int):
    """16-bit operand terms of a weight matrix as raw int16 bit patterns: split 3 = three bf16 terms, 2 = two fp16 terms."""
    if split == 3:
        return [t.view(torch.int16) for t in bf16_split3(wt)]
    if split == 2:
        return [t.view(torch.int16) for t in fp16_split2(wt)]  # third row block of the slab stays zero (unused)
    if split == 1:
        return [wt.to(torch.bfloat16).view(torch.int16)]       # bf16-storage arithmetic: one term, rows 2 and 3 unused
    raise ValueError(f"split must be 1 (one bf16 term), 2 (fp16 terms) or 3 (bf16 terms), got {split}")


def pack_tc3_weights(w: Tensor, kd: int, k: int, stride: int = 1, split: int = 3) -> Tensor:
    """[kd*k*k][Cin][Cout] fp32 -> the slab stream of the generation-3 tcgen05 kernel (conv_tc3.cu): one slab per
    (stage, tap) in the order of ``capi.conv_tc3_plan``; slab = [2 K-halves][w1 | w2 | w3, N rows each][8 bf16] over 16
    input channels (zero padded), N = max(Cout, 16).  ``split`` = 2: two fp16 terms [w1 | w2 | unused] for the
    MVSTER_TC3_FP16X2 arithmetic (same slab size).  Returned as a float32-typed byte blob."""
    from . import capi  # host-side plan enumeration lives in the library so that packer and kernel cannot diverge
    taps, cin, cout = w.shape
    if taps != kd * k * k:
        raise ValueError(f"weight has {taps} taps, expected kd*k*k = {kd * k * k}")
    n = 16 if cout < 16 else (80 if cout > 64 else cout)  # the UMMA N the kernel instantiates per weight split
    plan = capi.conv_tc3_plan(cin, kd, k, stride)
    out = torch.zeros((len(plan), 2, 3 * n, 8), dtype=torch.int16)
    w = w.detach().float().cpu()
    for i, (kz, ky, kx, c0, ky2, kx2) in enumerate(plan):
        wt = torch.zeros((16, n), dtype=torch.float32)  # rows = the MMA's K: 16 input channels, or (Cin <= 8) two taps x 8 channels
        if cin <= 8:
            wt[:cin, :cout] = w[(kz * k + ky) * k + kx, :, :]
            if ky2 >= 0:
                wt[8:8 + cin, :cout] = w[(kz * k + ky2) * k + kx2, :, :]
        else:
            cs = min(16, cin - c0)
            wt[:cs, :cout] = w[(kz * k + ky) * k + kx, c0:c0 + cs, :]
        for j, t in enumerate(_split_terms(wt, split)):
            for h in range(2):
                out[i, h, j * n:(j + 1) * n, :] = t[h * 8:(h + 1) * 8, :].T
    return out.reshape(-1).view(torch.float32)


def _tc3_slabs(mats, split: int = 3) -> Tensor:
    """list of [16][n] fp32 matrices (16 input channels x n accumulator columns) -> generation-3 slab stream."""
    n = mats[0].shape[1]
    out = torch.zeros((len(mats), 2, 3 * n, 8), dtype=torch.int16)
    for i, wt in enumerate(mats):
        for j, t in enumerate(_split_terms(wt, split)):
            for h in range(2):
                out[i, h, j * n:(j + 1) * n, :] = t[h * 8:(h + 1) * 8, :].T
    return out.reshape(-1).view(torch.float32)


def pack_tc3_deconv_weights(w: Tensor, rows: int = -1, split: int = 3) -> Tensor:
    """Transposed conv (1,3,3)/stride (1,2,2)/pad 1 weights [9][Cin][Cout] -> slab stream of mvster_deconv_tc3_f32: per
    16-channel chunk, one slab per input tap (dy,dx) in {0,1}^2 (rows = 0: dy = 0 only) whose columns are the output parity
    classes [class][Cout]; class (py,px) reads kernel element ky(py,dy), kx(px,dx) with k(0,0) = 1, k(1,0) = 2, k(1,1) = 0 and
    nothing for (0,1) (out[2i - 1 + k] += in[i] w[k])."""
    taps, cin, cout = w.shape
    if taps != 9 or cin % 16:
        raise ValueError(f"expected [9][Cin % 16 == 0][Cout], got {tuple(w.shape)}")
    w = w.detach().float().cpu()
    kidx = {(0, 0): 1, (1, 0): 2, (1, 1): 0}  # (output parity, input offset) -> kernel index
    classes = [(py, px) for py in ((0, 1) if rows < 0 else (rows,)) for px in (0, 1)]
    mats = []
    for c0 in range(0, cin, 16):
        for dy in ((0,) if rows == 0 else (0, 1)):
            for dx in (0, 1):
                m = torch.zeros((16, len(classes) * cout), dtype=torch.float32)
                for ci, (py, px) in enumerate(classes):
                    ky, kx = kidx.get((py, dy)), kidx.get((px, dx))
                    if ky is not None and kx is not None:
                        m[:, ci * cout:(ci + 1) * cout] = w[ky * 3 + kx, c0:c0 + 16, :]
                mats.append(m)
    return _tc3_slabs(mats, split)


def pack_reg3d(sd: Mapping[str, Tensor], prefix: str, layer_table) -> Tensor:
    """reg3d state (``{prefix}.conv0.conv.weight`` ... ``{prefix}.prob.weight``) -> blob in the layout of
    mvster_reg3d_layer_info (``layer_table`` = capi.reg3d_layer_table(G, down_size)); BN folded, prob has no bias."""
    last = layer_table[-1]
    blob = torch.zeros(last["w_off"] + 27 * last["cin"] * last["cout"], dtype=torch.float32)
    for L in layer_table:
        p = f"{prefix}.{L['name']}"
        if L["name"] == "prob":
            w = sd[p + ".weight"].detach().cpu().double()
            w = w.permute(2, 3, 4, 1, 0).reshape(27, L["cin"], 1).float().contiguous()
            b = None
        elif L["transposed"]:
            s, t = bn_scale_shift(sd, p + ".1")
            w, b = fold_deconv3d(sd[p + ".0.weight"].detach().cpu(), s.cpu(), t.cpu())
        else:
            s, t = bn_scale_shift(sd, p + ".bn")
            w, b = fold_conv3d(sd[p + ".conv.weight"].detach().cpu(), s.cpu(), t.cpu())
        if tuple(w.shape) != (27, L["cin"], L["cout"]):
            raise ValueError(f"{p}: packed shape {tuple(w.shape)} != {(27, L['cin'], L['cout'])}")
        blob[L["w_off"]:L["w_off"] + w.numel()] = w.reshape(-1)
        if b is not None:
            blob[L["b_off"]:L["b_off"] + b.numel()] = b
    return blob


REG2D_ORDER = ("conv0", "conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "conv7", "conv9", "conv11")


def pack_reg2d(sd: Mapping[str, Tensor], prefix: str, layer_table) -> Dict[str, Tensor]:
    """Pack one reg2d's state (keys ``{prefix}.conv0.conv.weight`` ...) into the blob layout
    reported by ``mvster_reg2d_layer_info`` (``layer_table`` = capi.reg2d_layer_table(G)).
    Returns {'blob': 1-D fp32, 'prob_w': [8], 'prob_b': [1]} on the CPU."""
    total = layer_table[-1]["b_off"] + layer_table[-1]["cout"]
    blob = torch.zeros(total, dtype=torch.float32)
    for name, L in zip(REG2D_ORDER, layer_table):
        p = f"{prefix}.{name}"
        if L["transposed"]:
            s, t = bn_scale_shift(sd, p + ".1")
            w, b = fold_deconv3d(sd[p + ".0.weight"].detach().cpu(), s.cpu(), t.cpu())
        else:
            s, t = bn_scale_shift(sd, p + ".bn")
            w, b = fold_conv3d(sd[p + ".conv.weight"].detach().cpu(), s.cpu(), t.cpu())
        if tuple(w.shape) != (L["taps"], L["cin"], L["cout"]):
            raise ValueError(f"{p}: packed shape {tuple(w.shape)} != layer table {(L['taps'], L['cin'], L['cout'])}")
        blob[L["w_off"]:L["w_off"] + w.numel()] = w.reshape(-1)
        blob[L["b_off"]:L["b_off"] + b.numel()] = b
    tc2 = []
    for name, L in zip(REG2D_ORDER, layer_table):
        if L["kd"] == 3:  # conv2 / conv4 / conv6: K-major [hi|lo] slabs for the tcgen05 paths
            w = blob[L["w_off"]:L["w_off"] + L["taps"] * L["cin"] * L["cout"]].reshape(L["taps"], L["cin"], L["cout"])
            tc2.append(pack_tc2_weights(w, 3))
    tc3 = {3: [], 2: []}  # three bf16 terms ('tc3_blob') and two fp16 terms ('tc3h_blob', MVSTER_TC3_FP16X2)
    for name, L in zip(REG2D_ORDER, layer_table):
        # slab streams of the generation-3 kernel in layer order (layout: conv_simt.cu tc3_layer_bytes)
        w = blob[L["w_off"]:L["w_off"] + L["taps"] * L["cin"] * L["cout"]].reshape(L["taps"], L["cin"], L["cout"])
        for split, lst in tc3.items():
            if not L["transposed"]:
                lst.append(pack_tc3_weights(w, L["kd"], 3, L["stride"], split))
            elif 4 * L["cout"] <= 64:   # all four output parity classes in one launch
                lst.append(pack_tc3_deconv_weights(w, -1, split))
            else:                       # conv7 (64 -> 32): output rows of parity 0, then parity 1
                lst += [pack_tc3_deconv_weights(w, 0, split), pack_tc3_deconv_weights(w, 1, split)]
    return {"blob": blob, "tc2_blob": torch.cat(tc2), "tc3_blob": torch.cat(tc3[3]),
            "tc3h_blob": torch.cat(tc3[2]),
            "prob_w": sd[prefix + ".prob.weight"].detach().cpu().reshape(-1).float().contiguous(),
            "prob_b": sd[prefix + ".prob.bias"].detach().cpu().reshape(-1).float().contiguous()}


def bf16_round(t: Tensor) -> Tensor:
    """fp32 values rounded to the nearest bf16 (ties to even), kept as fp32: what a bf16 store holds."""
    return t.float().to(torch.bfloat16).float()


def pack_reg2d_bf16(sd: Mapping[str, Tensor], prefix: str, layer_table) -> Dict[str, Tensor]:
    """bf16-storage arithmetic of one reg2d (mvster_reg2d_bf16): the convolution weights are rounded to bf16 UNFOLDED - the
    BatchNorm factor stays an fp32 per-channel scale applied to the accumulator, the shift goes to the bias slot - so that a
    layer computes BN(conv(bf16(x), bf16(w))) as the bf16 configuration is defined (oracle.storage).
    Returns {'blob_q', 'scales' [288], 'tc3_blob' (split = 1), 'prob_w' (bf16 values), 'prob_b'} on the CPU."""
    total = layer_table[-1]["b_off"] + layer_table[-1]["cout"]
    blob = torch.zeros(total, dtype=torch.float32)
    scales, tc3 = [], []
    one = lambda n: torch.ones(n, dtype=torch.float64)
    for name, L in zip(REG2D_ORDER, layer_table):
        p = f"{prefix}.{name}"
        if L["transposed"]:
            s, t = bn_scale_shift(sd, p + ".1")
            w, _ = fold_deconv3d(bf16_round(sd[p + ".0.weight"].detach().cpu()), one(L["cout"]), t.cpu())
        else:
            s, t = bn_scale_shift(sd, p + ".bn")
            w, _ = fold_conv3d(bf16_round(sd[p + ".conv.weight"].detach().cpu()), one(L["cout"]), t.cpu())
        blob[L["w_off"]:L["w_off"] + w.numel()] = w.reshape(-1)
        blob[L["b_off"]:L["b_off"] + L["cout"]] = t.float()
        scales.append(s.float().cpu())
        if not L["transposed"]:
            tc3.append(pack_tc3_weights(w, L["kd"], 3, L["stride"], 1))
        elif 4 * L["cout"] <= 64:
            tc3.append(pack_tc3_deconv_weights(w, -1, 1))
        else:
            tc3 += [pack_tc3_deconv_weights(w, 0, 1), pack_tc3_deconv_weights(w, 1, 1)]
    return {"blob_q": blob, "scales": torch.cat(scales).contiguous(), "tc3_blob": torch.cat(tc3),
            "prob_w": bf16_round(sd[prefix + ".prob.weight"].detach().cpu().reshape(-1)).contiguous(),
            "prob_b": sd[prefix + ".prob.bias"].detach().cpu().reshape(-1).float().contiguous()}
