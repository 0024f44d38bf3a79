"""CPU oracle for the MVSTER per-frame forward hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mvster_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker / the
reported CPU baseline.

What it is: a functional (no ``nn.Module``) restatement, in plain fp32 PyTorch
on the CPU, of the algorithm the reference executes in
``models/MVS4Net.py:60-111`` and ``models/mvs4net_utils.py`` (file:line cited
per function below).  Weights come in as a flat ``dict`` whose keys are the
reference's ``state_dict`` keys, so a reference checkpoint can be fed straight
in.

How it is pinned: the reference ships no tests, golden vectors or checkpoint
(SURVEY.md section 4), so the pin is the reference *itself*, imported from
``/root/reference`` in the build container by ``oracle/make_golden.py``; that
script runs the unmodified reference modules on seeded synthetic inputs and
stores inputs + outputs under ``tests/golden/``.  ``tests/test_oracle_golden.py``
replays those fixtures through this file (bit-exact or <=1e-6 where the
operation order of a fused ATen call is not contractually fixed).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
State = Dict[str, Tensor]

BN_EPS = 1e-5  # torch.nn.BatchNorm{2,3}d default, used by every BN in the reference


# --------------------------------------------------------------------------
# reduced-precision STORAGE (BASELINE cfg3 "bf16"): the reference has no bf16 path - its pixel grid is hard-coded fp32
# (mvs4net_utils.py:28-29) and torch.autocast destroys the geometry - so the bf16 configuration is
# defined as the fp32 reference with bf16 ROUNDING at the points where a bf16 build stores data: the FPN outputs, the cost
# volume, every convolution's operands (activations and weights) and the regulariser's stored layer outputs (skip sums included);
# accumulation, BN, geometry and softmax stay fp32.
# PARITY UNPINNED for this configuration: there is no reference output to pin it to; the fp32 path it is built on is pinned.
# ``fpn_internal=False`` is the variant the CUDA path implements (mvster_b200 storage="bf16"): the rounding applies to what the
# named hot path stores and reads - the pyramid's OUTPUT features, the cost volume, the regulariser's and the head's convolution
# operands - while the feature pyramid itself (outside the named path) computes in fp32.
# --------------------------------------------------------------------------
_STORAGE = None
_STORAGE_FPN_INTERNAL = True


class storage:
    """``with storage(torch.bfloat16): ...`` - emulate that storage type in every function below (None = exact fp32)."""

    def __init__(self, dtype, fpn_internal: bool = True):
        self.dtype = dtype
        self.fpn_internal = fpn_internal

    def __enter__(self):
        global _STORAGE, _STORAGE_FPN_INTERNAL
        self.prev, _STORAGE = (_STORAGE, _STORAGE_FPN_INTERNAL), self.dtype
        _STORAGE_FPN_INTERNAL = self.fpn_internal

    def __exit__(self, *exc):
        global _STORAGE, _STORAGE_FPN_INTERNAL
        _STORAGE, _STORAGE_FPN_INTERNAL = self.prev


def _q(x: Tensor) -> Tensor:
    return x if _STORAGE is None or not x.is_floating_point() else x.to(_STORAGE).to(x.dtype)


# --------------------------------------------------------------------------
# small building blocks
# --------------------------------------------------------------------------
def _bn_eval(x: Tensor, sd: State, p: str) -> Tensor:
    """Inference-mode batch norm from running statistics (reference modules are
    run under ``model.eval()``, test_mvs4.py:197)."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], False, 0.0, BN_EPS)


def _cbr3(x: Tensor, sd: State, p: str, stride=(1, 1, 1), pad=(1, 1, 1)) -> Tensor:
    """conv3d(no bias) -> BN -> ReLU, mvs4net_utils.py:116-123."""
    y = F.conv3d(_q(x), _q(sd[p + ".conv.weight"]), None, stride, pad)
    return F.relu(_bn_eval(y, sd, p + ".bn"))


def _up3(x: Tensor, sd: State, p: str, stride, pad, out_pad) -> Tensor:
    """ConvTranspose3d(no bias) -> BN -> ReLU packed as nn.Sequential
    (indices .0/.1), mvs4net_utils.py:885-898 and :926-940."""
    y = F.conv_transpose3d(_q(x), _q(sd[p + ".0.weight"]), None, stride, pad, out_pad)
    return F.relu(_bn_eval(y, sd, p + ".1"))


def _cbr2(x: Tensor, sd: State, p: str, stride: int, pad: int, relu: bool = True) -> Tensor:
    """2-D conv -> BN -> ReLU block of the feature net, mvs4net_utils.py:224-251."""
    y = _bn_eval(F.conv2d(_q(x), _q(sd[p + ".conv.weight"]), None, stride, pad), sd, p + ".bn")
    return F.relu(y) if relu else y


# --------------------------------------------------------------------------
# FPN4 feature pyramid (outside the named hot path, needed to drive it)
# --------------------------------------------------------------------------
def fpn4_features(sd: State, img: Tensor, p: str = "feature") -> Dict[str, Tensor]:
    """mvs4net_utils.py:472-502.  img [B,3,H,W] -> stage1..4 features with
    64/32/16/8 channels at H/8 .. H."""
    if _STORAGE is not None and not _STORAGE_FPN_INTERNAL:  # fp32 pyramid, reduced-precision storage of its outputs only
        dtype = _STORAGE
        with storage(None):
            full = fpn4_features(sd, img, p)
        with storage(dtype, False):
            return {k: _q(v) for k, v in full.items()}
    c0 = _cbr2(_cbr2(img, sd, p + ".conv0.0", 1, 1), sd, p + ".conv0.1", 1, 1)
    c1 = _cbr2(c0, sd, p + ".conv1.0", 2, 2)
    c1 = _cbr2(_cbr2(c1, sd, p + ".conv1.1", 1, 1), sd, p + ".conv1.2", 1, 1)
    c2 = _cbr2(c1, sd, p + ".conv2.0", 2, 2)
    c2 = _cbr2(_cbr2(c2, sd, p + ".conv2.1", 1, 1), sd, p + ".conv2.2", 1, 1)
    c3 = _cbr2(c2, sd, p + ".conv3.0", 2, 2)
    c3 = _cbr2(_cbr2(c3, sd, p + ".conv3.1", 1, 1), sd, p + ".conv3.2", 1, 1)

    def up(t):
        return F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=True)

    def w(name):
        return _q(sd[p + name])

    out = {}
    top = c3
    out["stage1"] = _q(F.conv2d(_q(top), w(".out1.weight")))
    top = up(top) + F.conv2d(_q(c2), w(".inner1.weight"), sd[p + ".inner1.bias"])
    out["stage2"] = _q(F.conv2d(_q(top), w(".out2.weight"), None, 1, 1))
    top = up(top) + F.conv2d(_q(c1), w(".inner2.weight"), sd[p + ".inner2.bias"])
    out["stage3"] = _q(F.conv2d(_q(top), w(".out3.weight"), None, 1, 1))
    top = up(top) + F.conv2d(_q(c0), w(".inner3.weight"), sd[p + ".inner3.bias"])
    out["stage4"] = _q(F.conv2d(_q(top), w(".out4.weight"), None, 1, 1))
    return out


# --------------------------------------------------------------------------
# depth hypothesis schedulers
# --------------------------------------------------------------------------
def hypo_init_inverse(depth_values: Tensor, D: int, H: int, W: int) -> Tensor:
    """mvs4net_utils.py:71-77: uniform in 1/d between 1/d_max and 1/d_min,
    ordered far -> near, constant over pixels.  Returns [B,D,H,W]."""
    inv_near = 1.0 / depth_values[:, 0]
    inv_far = 1.0 / depth_values[:, -1]
    ramp = torch.arange(0, D, dtype=depth_values.dtype).reshape(1, -1, 1, 1).repeat(1, 1, H, W) / (D - 1)
    inv = inv_far[:, None, None, None] + (inv_near - inv_far)[:, None, None, None] * ramp
    return 1.0 / inv


def hypo_schedule_inverse(inv_min: Tensor, inv_max: Tensor, D: int, H: int, W: int) -> Tensor:
    """mvs4net_utils.py:79-86: per coarse pixel D samples between inv_max and
    inv_min, trilinear (align_corners) up to (D,H,W), reciprocal."""
    ramp = torch.arange(0, D, dtype=inv_min.dtype).reshape(1, -1, 1, 1).repeat(1, 1, H // 2, W // 2) / (D - 1)
    inv = inv_max[:, None] + (inv_min - inv_max)[:, None] * ramp
    inv = F.interpolate(inv.unsqueeze(1), [D, H, W], mode="trilinear", align_corners=True).squeeze(1)
    return 1.0 / inv


def hypo_init_linear(depth_values: Tensor, D: int, H: int, W: int) -> Tensor:
    """mvs4net_utils.py:61-69 (non-inverse option)."""
    d0 = depth_values[:, 0]
    step = (depth_values[:, -1] - d0) / (D - 1)
    s = d0.unsqueeze(1) + torch.arange(0, D, dtype=depth_values.dtype).reshape(1, -1) * step[:, None]
    return s[:, :, None, None].repeat(1, 1, H, W)


def hypo_schedule_linear(depth: Tensor, D: int, itv_pixel: Tensor, H: int, W: int) -> Tensor:
    """mvs4net_utils.py:88-99 (non-inverse option); itv_pixel is [B]."""
    lo = depth - D / 2 * itv_pixel[:, None, None]
    hi = depth + D / 2 * itv_pixel[:, None, None]
    step = (hi - lo) / (D - 1)
    s = lo.unsqueeze(1) + torch.arange(0, D, dtype=depth.dtype).reshape(1, -1, 1, 1) * step.unsqueeze(1)
    return F.interpolate(s.unsqueeze(1), [D, H, W], mode="trilinear", align_corners=True).squeeze(1)


# --------------------------------------------------------------------------
# geometry + warp
# --------------------------------------------------------------------------
def compose_projection(cam: Tensor) -> Tensor:
    """cam [B,2,4,4] (extrinsic, intrinsic) -> full 4x4 whose top 3x4 block is
    K @ E[:3,:4]; mvs4net_utils.py:1032-1035."""
    full = cam[:, 0].clone()
    full[:, :3, :4] = torch.matmul(cam[:, 1, :3, :3], cam[:, 0, :3, :4])
    return full


def relative_pose(src_full: Tensor, ref_full: Tensor):
    """mvs4net_utils.py:24-26: P = src @ inv(ref); returns (R [B,3,3], t [B,3,1])."""
    P = torch.matmul(src_full, torch.inverse(ref_full))
    return P[:, :3, :3], P[:, :3, 3:4]


def plane_sweep_warp(src_fea: Tensor, src_full: Tensor, ref_full: Tensor, hypo: Tensor) -> Tensor:
    """mvs4net_utils.py:13-59.  src_fea [B,C,Hs,Ws], hypo [B,D,Hr,Wr] ->
    [B,C,D,Hr,Wr]; bilinear, zero padding per tap, align_corners=True; z==0 is
    replaced by 1e-9; the reference's pixel grid is hard-coded fp32 (:28-29).  Feeding fp64
    tensors runs the whole chain in fp64 (the reference cannot) - used only as the
    higher-precision "truth" that brackets fp32 rounding noise in the GPU parity tests."""
    B, C, Hs, Ws = src_fea.shape
    _, D, Hr, Wr = hypo.shape
    R, t = relative_pose(src_full, ref_full)
    gdt = torch.float64 if src_fea.dtype == torch.float64 else torch.float32
    yy, xx = torch.meshgrid(torch.arange(0, Hr, dtype=gdt),
                            torch.arange(0, Wr, dtype=gdt), indexing="ij")
    pix = torch.stack((xx.reshape(-1), yy.reshape(-1), torch.ones(Hr * Wr, dtype=gdt)))  # [3,HW]
    ray = torch.matmul(R, pix.unsqueeze(0).repeat(B, 1, 1))                   # [B,3,HW]
    pts = ray.unsqueeze(2).repeat(1, 1, D, 1) * hypo.reshape(B, 1, D, -1) + t.reshape(B, 3, 1, 1)
    z = pts[:, 2:3]
    z[z == 0] = 1e-9  # in-place on a view of pts, as the reference does
    uv = pts[:, :2] / z
    gx = uv[:, 0] / ((Ws - 1) / 2) - 1
    gy = uv[:, 1] / ((Hs - 1) / 2) - 1
    grid = torch.stack((gx, gy), dim=3)
    out = F.grid_sample(src_fea, grid.reshape(B, D * Hr, Wr, 2), mode="bilinear",
                        padding_mode="zeros", align_corners=True)
    return out.reshape(B, C, D, Hr, Wr)


# --------------------------------------------------------------------------
# Epipolar-Transformer aggregation
# --------------------------------------------------------------------------
def et_view_terms(ref_fea: Tensor, src_fea: Tensor, ref_cam: Tensor, src_cam: Tensor, hypo: Tensor,
                  group_cor: bool, G: int, attn_temp: float, attn_fuse_d: bool = True):
    """One source view's contribution, mvs4net_utils.py:1032-1055.
    Returns (w * cor [B,G',D,H,W], w) with w [B,D,H,W] (or [B,H,W] when
    attn_fuse_d is False).  G' = G with group correlation, C otherwise."""
    B, D, H, W = hypo.shape
    C = ref_fea.shape[1]
    warped = plane_sweep_warp(src_fea, compose_projection(src_cam), compose_projection(ref_cam), hypo)
    ref_vol = ref_fea.unsqueeze(2).repeat(1, 1, D, 1, 1)
    if group_cor:
        cor = (warped.reshape(B, G, C // G, D, H, W) * ref_vol.reshape(B, G, C // G, D, H, W)).mean(2)
    else:
        cor = (ref_vol - warped) ** 2
    if attn_fuse_d:
        w = torch.softmax(cor.sum(1) / attn_temp, 1) / math.sqrt(C)
        return w.unsqueeze(1) * cor, w
    w = torch.softmax(cor.sum(1), 1).max(1)[0]
    return w.unsqueeze(1).unsqueeze(1) * cor, w


def et_aggregate(features: Sequence[Tensor], cams: Tensor, hypo: Tensor, group_cor: bool, G: int,
                 attn_temp: float, attn_fuse_d: bool = True,
                 views: Optional[Sequence[int]] = None, partial: bool = False):
    """mvs4net_utils.py:1015-1062.  features: [ref, src1, ...] each [B,C,H,W];
    cams [B,Nv,2,4,4].  ``views`` restricts to a subset of source-view indices
    (1-based positions in ``features``) and ``partial=True`` returns the
    un-normalised (acc, wsum) pair without the 1e-8 seed - the quantities a
    view-sharded run all-reduces."""
    idx = list(range(1, len(features))) if views is None else list(views)
    acc = 0
    wsum = 0.0 if partial else 1e-8
    for v in idx:
        term, w = et_view_terms(features[0], features[v], cams[:, 0], cams[:, v], hypo,
                                group_cor, G, attn_temp, attn_fuse_d)
        wsum = wsum + w
        acc = acc + term
    if partial:
        return acc, wsum
    if attn_fuse_d:
        return acc / wsum.unsqueeze(1)
    return acc / wsum.unsqueeze(1).unsqueeze(1)


def et_normalize(acc: Tensor, wsum: Tensor) -> Tensor:
    """Division of all-reduced partials; equals mvs4net_utils.py:1060 with the
    1e-8 seed of :1022 added once."""
    return acc / (1e-8 + wsum).unsqueeze(1)


# --------------------------------------------------------------------------
# regularisation networks
# --------------------------------------------------------------------------
def reg2d_logits(sd: State, p: str, cost: Tensor) -> Tensor:
    """mvs4net_utils.py:870-912.  cost [B,G,D,H,W] -> logits [B,D,H,W]."""
    s2, p2 = (1, 2, 2), (0, 1, 1)
    # _q(...) marks what a reduced-precision build STORES (identity in fp32): every layer's output, skip sums included; the last
    # sum feeds the `prob` layer, which rounds its operand itself
    c0 = _q(_cbr3(cost, sd, p + ".conv0", (1, 1, 1), p2))
    c2 = _q(_cbr3(_q(_cbr3(c0, sd, p + ".conv1", s2, p2)), sd, p + ".conv2"))
    c4 = _q(_cbr3(_q(_cbr3(c2, sd, p + ".conv3", s2, p2)), sd, p + ".conv4"))
    x = _q(_cbr3(_q(_cbr3(c4, sd, p + ".conv5", s2, p2)), sd, p + ".conv6"))
    x = _q(c4 + _up3(x, sd, p + ".conv7", s2, p2, p2))
    x = _q(c2 + _up3(x, sd, p + ".conv9", s2, p2, p2))
    x = c0 + _up3(x, sd, p + ".conv11", s2, p2, p2)
    return F.conv3d(_q(x), _q(sd[p + ".prob.weight"]), sd[p + ".prob.bias"]).squeeze(1)


def reg3d_logits(sd: State, p: str, cost: Tensor, down_size: int) -> Tensor:
    """mvs4net_utils.py:914-965: full 3-D U-Net, stride 2 in D as well."""
    s2, one = (2, 2, 2), (1, 1, 1)
    c0 = _cbr3(cost, sd, p + ".conv0")
    c2 = _cbr3(_cbr3(c0, sd, p + ".conv1", s2, one), sd, p + ".conv2")
    x = c2
    if down_size >= 2:
        c4 = _cbr3(_cbr3(c2, sd, p + ".conv3", s2, one), sd, p + ".conv4")
        x = c4
        if down_size >= 3:
            x = _cbr3(_cbr3(c4, sd, p + ".conv5", s2, one), sd, p + ".conv6")
            x = c4 + _up3(x, sd, p + ".conv7", s2, one, one)
        x = c2 + _up3(x, sd, p + ".conv9", s2, one, one)
    x = c0 + _up3(x, sd, p + ".conv11", s2, one, one)
    return F.conv3d(_q(x), _q(sd[p + ".prob.weight"]), None, 1, 1).squeeze(1)


# --------------------------------------------------------------------------
# head
# --------------------------------------------------------------------------
def depth_head(logits: Tensor, hypo: Tensor, stage_idx: int, split_itv: float, inverse_depth: bool,
               full_res_conf: bool = True) -> Dict[str, Tensor]:
    """mvs4net_utils.py:1066-1088: softmax over D, winner-take-all depth,
    eval-mode confidence (max prob, bilinear x2^(3-stage) align_corners),
    next-stage inverse range."""
    attn = F.softmax(logits, dim=1)
    best = attn.max(1, keepdim=True)[1]
    depth = torch.gather(hypo, 1, best).squeeze(1)
    conf = attn.max(1)[0]
    if full_res_conf:
        conf = F.interpolate(conf.unsqueeze(1), scale_factor=2 ** (3 - stage_idx), mode="bilinear",
                             align_corners=True).squeeze(1)
    out = {"depth": depth, "photometric_confidence": conf, "hypo_depth": hypo, "attn_weight": attn}
    if inverse_depth:
        itv = 1.0 / hypo[:, 2] - 1.0 / hypo[:, 1]
        out["inverse_min_depth"] = 1 / depth + split_itv * itv
        out["inverse_max_depth"] = 1 / depth - split_itv * itv
    return out


def soft_depth_regression(prob: Tensor, depth_values: Tensor) -> Tensor:
    """models/module.py:935-941 (dead code in the reference, optional output
    here): expectation sum_d p*d."""
    if depth_values.dim() <= 2:
        depth_values = depth_values.view(*depth_values.shape, 1, 1)
    return torch.sum(prob * depth_values, 1)


# --------------------------------------------------------------------------
# whole cascade
# --------------------------------------------------------------------------
DEFAULT_CFG = dict(reg_net="reg2d", num_stage=4, stage_splits=[8, 8, 4, 4],
                   depth_interals_ratio=[0.5, 0.5, 0.5, 1], group_cor=True,
                   group_cor_dim=[8, 8, 4, 4], inverse_depth=True, attn_temp=2.0,
                   attn_fuse_d=True, mono=True)
REG3D_DOWN = [3, 3, 2, 2]  # MVS4Net.py:48


def stage_forward(sd: State, cfg: dict, k: int, feats: List[Tensor], cams: Tensor, hypo: Tensor) -> Dict[str, Tensor]:
    """stagenet.forward for stage index k (0-based), mvs4net_utils.py:1012-1094 (eval mode)."""
    G = cfg["group_cor_dim"][k]
    cost = _q(et_aggregate(feats, cams, hypo, cfg["group_cor"], G, cfg["attn_temp"], cfg.get("attn_fuse_d", True)))
    if cfg["reg_net"] == "reg2d":
        logits = reg2d_logits(sd, f"reg.{k}", cost)
    else:
        logits = reg3d_logits(sd, f"reg.{k}", cost, REG3D_DOWN[k])
    out = depth_head(logits, hypo, k, cfg["depth_interals_ratio"][k], cfg["inverse_depth"])
    out["cost"] = cost      # extra (not in the reference dict): for teacher-forced checks
    out["logits"] = logits  # extra
    if cfg.get("mono", False):
        out["mono_feat"] = feats[0]
    return out


def cascade_forward(sd: State, cfg: dict, imgs: Sequence[Tensor], proj_matrices: Dict[str, Tensor],
                    depth_values: Tensor, features: Optional[List[Dict[str, Tensor]]] = None, storage_dtype=None,
                    storage_fpn_internal: bool = True) -> Dict:
    """MVS4net.forward in eval mode, MVS4Net.py:60-111.  ``storage_dtype=torch.bfloat16`` evaluates the bf16-storage
    configuration (see ``storage`` above; ``storage_fpn_internal=False``: the variant the CUDA path implements); None is the
    reference's fp32."""
    with torch.no_grad(), storage(storage_dtype, storage_fpn_internal):
        if features is None:
            features = [fpn4_features(sd, im) for im in imgs]
        outputs: Dict = {}
        prev = None
        depth_itv = (depth_values[:, -1] - depth_values[:, 0]) / depth_values.size(1)
        for k in range(cfg["num_stage"]):
            key = f"stage{k + 1}"
            feats = [f[key] for f in features]
            B, C, H, W = feats[0].shape
            D = cfg["stage_splits"][k]
            if k == 0:
                hypo = hypo_init_inverse(depth_values, D, H, W) if cfg["inverse_depth"] \
                    else hypo_init_linear(depth_values, D, H, W)
            elif cfg["inverse_depth"]:
                hypo = hypo_schedule_inverse(prev["inverse_min_depth"], prev["inverse_max_depth"], D, H, W)
            else:
                hypo = hypo_schedule_linear(prev["depth"], D, cfg["depth_interals_ratio"][k] * depth_itv, H, W)
            prev = stage_forward(sd, cfg, k, feats, proj_matrices[key], hypo)
            outputs[key] = prev
            outputs.update(prev)
        return outputs
