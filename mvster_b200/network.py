"""``MVS4net`` - drop-in mirror of the reference module API (models/MVS4Net.py:9-111).

Same constructor keywords, same sub-module attribute names (=> identical ``state_dict``
keys: ``feature.*``, ``reg.{k}.*``, ``mono_depth_decoder.*``), same ``forward(imgs,
proj_matrices, depth_values, filename=None) -> dict`` contract, same output keys.

Dispatch in ``forward``:
  * eval mode under ``torch.no_grad()`` on CUDA tensors  -> the sm_100a library through
    the C ABI (engine.py).  This is the product path; it raises if the library is missing.
  * training / grad enabled -> differentiable PyTorch ops (torch_path.py), any device; with ``MVSTER_TRAIN_ET=1`` (or
    ``model.stagenet.train_et = True``) the warp + ET aggregation of CUDA features runs on the fused kernel and its
    hand-written backward (train_ops.py, csrc/et_fuse_bwd.cu) instead of materialising the warped volumes.
  * eval + no_grad on CPU tensors -> error (no CPU fallback by design).
"""
from __future__ import annotations

import os
from typing import Dict

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import torch_path as tp
from . import train_ops

Tensor = torch.Tensor


# ------------------------------------------------------------------ building blocks
class ConvBn2d(nn.Module):
    """conv -> BN -> ReLU with children named ``conv`` / ``bn`` (mvs4net_utils.py:224-251)."""

    def __init__(self, cin, cout, k, stride=1, padding=0):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride=stride, padding=padding, bias=False)
        self.bn = nn.BatchNorm2d(cout)

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)))


class ConvBn3d(nn.Module):
    """mvs4net_utils.py:116-123."""

    def __init__(self, cin, cout, k=3, stride=1, pad=1):
        super().__init__()
        self.conv = nn.Conv3d(cin, cout, k, stride=stride, padding=pad, bias=False)
        self.bn = nn.BatchNorm3d(cout)

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)))


def _up3d(cin, cout, k, stride, pad, out_pad):
    return nn.Sequential(nn.ConvTranspose3d(cin, cout, k, stride=stride, padding=pad, output_padding=out_pad, bias=False),
                         nn.BatchNorm3d(cout), nn.ReLU(inplace=True))


class FPN4(nn.Module):
    """4-scale feature pyramid, mvs4net_utils.py:419-502 (gn=False, dcn=False)."""

    def __init__(self, base_channels: int = 8):
        super().__init__()
        b = base_channels
        self.base_channels = b

        def level(cin, cout):
            return nn.Sequential(ConvBn2d(cin, cout, 5, 2, 2), ConvBn2d(cout, cout, 3, 1, 1), ConvBn2d(cout, cout, 3, 1, 1))

        self.conv0 = nn.Sequential(ConvBn2d(3, b, 3, 1, 1), ConvBn2d(b, b, 3, 1, 1))
        self.conv1 = level(b, 2 * b)
        self.conv2 = level(2 * b, 4 * b)
        self.conv3 = level(4 * b, 8 * b)
        top = 8 * b
        self.inner1 = nn.Conv2d(4 * b, top, 1, bias=True)
        self.inner2 = nn.Conv2d(2 * b, top, 1, bias=True)
        self.inner3 = nn.Conv2d(b, top, 1, bias=True)
        self.out1 = nn.Conv2d(top, 8 * b, 1, bias=False)
        self.out2 = nn.Conv2d(top, 4 * b, 3, padding=1, bias=False)
        self.out3 = nn.Conv2d(top, 2 * b, 3, padding=1, bias=False)
        self.out4 = nn.Conv2d(top, b, 3, padding=1, bias=False)
        self.out_channels = [8 * b, 4 * b, 2 * b, b]

    def forward(self, x: Tensor) -> Dict[str, Tensor]:
        c0 = self.conv0(x)
        c1 = self.conv1(c0)
        c2 = self.conv2(c1)
        c3 = self.conv3(c2)
        up = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=True)
        out = {"stage1": self.out1(c3)}
        t = up(c3) + self.inner1(c2)
        out["stage2"] = self.out2(t)
        t = up(t) + self.inner2(c1)
        out["stage3"] = self.out3(t)
        t = up(t) + self.inner3(c0)
        out["stage4"] = self.out4(t)
        return out


class reg2d(nn.Module):
    """mvs4net_utils.py:870-912 (conv_name == 'ConvBnReLU3D')."""

    def __init__(self, input_channel=128, base_channel=32, conv_name="ConvBnReLU3D"):
        super().__init__()
        if conv_name != "ConvBnReLU3D":
            raise NotImplementedError(f"agg_type={conv_name!r}: only 'ConvBnReLU3D' (the shipped config) is implemented")
        b, k, p, s = base_channel, (1, 3, 3), (0, 1, 1), (1, 2, 2)
        self.conv0 = ConvBn3d(input_channel, b, k, 1, p)
        self.conv1 = ConvBn3d(b, 2 * b, k, s, p)
        self.conv2 = ConvBn3d(2 * b, 2 * b)
        self.conv3 = ConvBn3d(2 * b, 4 * b, k, s, p)
        self.conv4 = ConvBn3d(4 * b, 4 * b)
        self.conv5 = ConvBn3d(4 * b, 8 * b, k, s, p)
        self.conv6 = ConvBn3d(8 * b, 8 * b)
        self.conv7 = _up3d(8 * b, 4 * b, k, s, p, p)
        self.conv9 = _up3d(4 * b, 2 * b, k, s, p, p)
        self.conv11 = _up3d(2 * b, b, k, s, p, p)
        self.prob = nn.Conv3d(8, 1, 1, stride=1, padding=0)  # 8 input channels hard-coded, :900

    def forward(self, x):
        c0 = self.conv0(x)
        c2 = self.conv2(self.conv1(c0))
        c4 = self.conv4(self.conv3(c2))
        x = self.conv6(self.conv5(c4))
        x = c4 + self.conv7(x)
        x = c2 + self.conv9(x)
        x = c0 + self.conv11(x)
        return self.prob(x).squeeze(1)


class reg3d(nn.Module):
    """mvs4net_utils.py:914-965."""

    def __init__(self, in_channels, base_channels, down_size=3):
        super().__init__()
        b = base_channels
        self.down_size = down_size
        self.conv0 = ConvBn3d(in_channels, b)
        self.conv1 = ConvBn3d(b, 2 * b, 3, 2, 1)
        self.conv2 = ConvBn3d(2 * b, 2 * b)
        if down_size >= 2:
            self.conv3 = ConvBn3d(2 * b, 4 * b, 3, 2, 1)
            self.conv4 = ConvBn3d(4 * b, 4 * b)
        if down_size >= 3:
            self.conv5 = ConvBn3d(4 * b, 8 * b, 3, 2, 1)
            self.conv6 = ConvBn3d(8 * b, 8 * b)
            self.conv7 = _up3d(8 * b, 4 * b, 3, 2, 1, 1)
        if down_size >= 2:
            self.conv9 = _up3d(4 * b, 2 * b, 3, 2, 1, 1)
        self.conv11 = _up3d(2 * b, b, 3, 2, 1, 1)
        self.prob = nn.Conv3d(b, 1, 3, stride=1, padding=1, bias=False)

    def forward(self, x):
        c0 = self.conv0(x)
        c2 = self.conv2(self.conv1(c0))
        x = c2
        if self.down_size >= 2:
            c4 = self.conv4(self.conv3(c2))
            x = c4
            if self.down_size >= 3:
                x = c4 + self.conv7(self.conv6(self.conv5(c4)))
            x = c2 + self.conv9(x)
        x = c0 + self.conv11(x)
        return self.prob(x).squeeze(1)


class mono_depth_decoder(nn.Module):
    """Train-only monocular auxiliary head, mvs4net_utils.py:833-868."""

    def __init__(self):
        super().__init__()
        self.convblocks = nn.ModuleList([ConvBn2d(64, 32, 3, 1, 1), ConvBn2d(32, 16, 3, 1, 1), ConvBn2d(16, 8, 3, 1, 1)])
        self.conv3x3 = nn.ModuleList([nn.Conv2d(64, 1, 3, 1, 1), nn.Conv2d(32, 1, 3, 1, 1), nn.Conv2d(16, 1, 3, 1, 1)])

    def forward(self, outputs, d_min, d_max):
        lo, hi = (1 / d_max)[:, None, None, None], (1 / d_min)[:, None, None, None]
        for i in range(1, 4):
            small = self.convblocks[i - 1](outputs[f"stage{i}"]["mono_feat"])
            small = F.interpolate(small, scale_factor=2, mode="nearest")
            disp = torch.sigmoid(self.conv3x3[i - 1](torch.cat([small, outputs[f"stage{i + 1}"]["mono_feat"]], 1)))
            outputs[f"stage{i + 1}"]["mono_depth"] = (1 / (lo + (hi - lo) * disp)).squeeze(1)
        return outputs


class stagenet(nn.Module):
    """Parameter-less stage operator (mvs4net_utils.py:1003-1094).  ``forward`` is the
    differentiable PyTorch formulation; the CUDA inference path lives in engine.py and is
    selected by ``MVS4net.forward``."""

    def __init__(self, inverse_depth=False, mono=False, attn_fuse_d=True, vis_ETA=False, attn_temp=1):
        super().__init__()
        self.inverse_depth, self.mono, self.attn_fuse_d = inverse_depth, mono, attn_fuse_d
        self.vis_ETA, self.attn_temp = vis_ETA, attn_temp
        self.train_et = None  # True / False: aggregate CUDA features through the fused kernels + their backward; None: MVSTER_TRAIN_ET

    def forward(self, features, proj_matrices, depth_hypo, regnet, stage_idx, group_cor=False, group_cor_dim=8,
                split_itv=1, fn=None):
        if self.vis_ETA:
            raise NotImplementedError("vis_ETA debug dumps are not part of this implementation")
        kernels = train_ops.enabled_by_default() if self.train_et is None else self.train_et
        if kernels and group_cor and self.attn_fuse_d and train_ops.usable(features, group_cor_dim, depth_hypo.shape[1]):
            cost = train_ops.aggregate(features, proj_matrices, depth_hypo, group_cor_dim, self.attn_temp)
        else:
            cost = tp.aggregate(features, proj_matrices, depth_hypo, group_cor, group_cor_dim, self.attn_temp, self.attn_fuse_d)
        out = tp.head(regnet(cost), depth_hypo, stage_idx, split_itv, self.inverse_depth, self.training)
        if self.mono:
            out["mono_feat"] = features[0]
        return out


# ------------------------------------------------------------------ the model
class MVS4net(nn.Module):
    def __init__(self, arch_mode="fpn", reg_net="reg2d", num_stage=4, fpn_base_channel=8, reg_channel=8,
                 stage_splits=[8, 8, 4, 4], depth_interals_ratio=[0.5, 0.5, 0.5, 1], group_cor=False,
                 group_cor_dim=[8, 8, 8, 8], inverse_depth=False, agg_type="ConvBnReLU3D", dcn=False, pos_enc=0,
                 mono=False, asff=False, attn_temp=2, attn_fuse_d=True, vis_ETA=False, vis_mono=False):
        super().__init__()
        if arch_mode != "fpn":
            raise NotImplementedError("arch_mode must be 'fpn' (the only feature net the reference constructs)")
        if dcn:
            raise NotImplementedError("dcn=True needs an external, un-vendored deformable-conv extension (mvs4net_utils.py:5-9)")
        if asff:
            raise NotImplementedError("asff=True (ablation branch, not in the shipped config) is not implemented")
        if vis_mono:
            raise NotImplementedError("vis_mono debug dumps are not implemented")
        self.arch_mode, self.num_stage = arch_mode, num_stage
        self.depth_interals_ratio = list(depth_interals_ratio)
        self.group_cor, self.group_cor_dim = group_cor, list(group_cor_dim)
        self.inverse_depth, self.asff = inverse_depth, asff
        self.attn_ob = nn.ModuleList()
        self.feature = FPN4(base_channels=fpn_base_channel)
        self.vis_mono = vis_mono
        self.stagenet = stagenet(inverse_depth, mono, attn_fuse_d, vis_ETA, attn_temp)
        self.stage_splits = list(stage_splits)
        self.reg = nn.ModuleList()
        self.pos_enc = pos_enc
        self.pos_enc_func = nn.ModuleList()
        self.mono = mono
        self.reg_net = reg_net
        if mono:
            self.mono_depth_decoder = mono_depth_decoder()
        if reg_net == "reg3d":
            self.down_size = [3, 3, 2, 2]
        for k in range(num_stage):
            cin = self.group_cor_dim[k] if group_cor else self.feature.out_channels[k]
            if reg_net == "reg2d":
                self.reg.append(reg2d(input_channel=cin, base_channel=reg_channel, conv_name=agg_type))
            elif reg_net == "reg3d":
                self.reg.append(reg3d(in_channels=cin, base_channels=reg_channel, down_size=self.down_size[k]))
            else:
                raise ValueError(f"unknown reg_net {reg_net!r}")
        # Folded/packed device weights live in one InferenceEngine per CUDA device (nn.DataParallel
        # replicas share this dict by reference, each thread touching only its own device's entry).
        self._engines = {}
        self._weights_version = 0
        # arithmetic of the regulariser convolutions on the CUDA inference path:
        #   "fp32"   every layer on the CUDA cores (exact fp32 FMA)
        #   "2xfp16" conv1..conv11 on the persistent tcgen05 kernel (conv_tc3.cu): two fp16 terms per operand (22-bit operands,
        #            per layer <= 2.1e-6 of max against an fp64 convolution, tests/test_gpu_tc_conv.py) - the default; valid while
        #            the layer inputs stay below 65504 in magnitude (they are BN-normalised activations; larger values saturate)
        #   "3xbf16" same kernel, three bf16 terms per operand (24-bit operands, full fp32 range; <= 5e-6 of max), 1.5x the MMAs
        #   "3xtf32" the 3x3x3 layers on the staged-tile TF32 kernel (conv_tc2.cu), error-compensated TF32 (<= 1.6e-5 of max)
        #   "tf32"   ... single TF32 pass (reduced precision, opt-in)
        self.reg_precision = os.environ.get("MVSTER_REG_PRECISION", "2xfp16")
        # storage type of the hot path's tensors at inference: "fp32" (the reference's), or "bf16" (BASELINE configs[2]; the
        # reference has no such path): the pyramid's output features and the cost volume are bf16 in HBM and the regulariser /
        # head convolutions take bf16 operands (one-term tcgen05 arithmetic, BatchNorm scale in fp32); geometry, correlations,
        # softmax and accumulation stay fp32; the feature pyramid itself computes with fpn_precision.  reg2d + group correlation
        # + attn_fuse_d only (the shipped configuration); no view sharding.
        self.storage = os.environ.get("MVSTER_STORAGE", "fp32")
        # feature pyramid at inference: "native" (libmvster_b200 kernels, fpn_engine.py; fpn_precision as above, "3xbf16" puts
        # every layer after the 3-channel stem on the tensor cores) or "torch" (the module's own convs through cuDNN, channels-last)
        self.fpn_backend = os.environ.get("MVSTER_FPN", "native")
        self.fpn_precision = os.environ.get("MVSTER_FPN_PRECISION", "2xfp16")
        # run cascade stages 1-3 on a second stream next to the feature pyramid's last levels (engine._forward_overlapped)
        self.overlap_stages = os.environ.get("MVSTER_OVERLAP", "1") == "1"
        # replay the whole inference forward as one CUDA graph per input signature (engine.forward_graphed; MVSTER_CUDA_GRAPH=0 =
        # eager launches).  Results are returned as copies unless graph_static_outputs is set (then they are the graph's static
        # buffers, valid until the next call with the same signature).
        self.use_cuda_graph = os.environ.get("MVSTER_CUDA_GRAPH", "1") == "1"
        self.graph_static_outputs = os.environ.get("MVSTER_GRAPH_STATIC_OUTPUTS", "0") == "1"
        # several frames in flight (experimental): each value of graph_slot owns its own captured graph (static inputs, workspace,
        # outputs), so a serving loop can replay slot 0 and slot 1 on two streams.  Measured with bench.py's MVSTER_BENCH_INFLIGHT=2:
        # +15-19 % in some runs, 2-4x slower in others (both graphs' persistent kernels size their grids for the whole GPU)
        self.graph_slot = 0
        self._view_shard = None  # sharding.ViewShard: this rank's slice of the source views (multi-GPU inference)

    def set_view_shard(self, shard) -> None:
        """View-parallel inference: this process aggregates only ``shard.views`` and all-reduces the
        per-stage partials with the other ranks of ``shard.group`` (mvster_b200/sharding.py)."""
        self._view_shard = shard

    # -- weight-change tracking: packed weights are rebuilt when parameters may have moved
    def train(self, mode: bool = True):
        self._weights_version += 1
        return super().train(mode)

    def load_state_dict(self, *a, **kw):
        self._weights_version += 1
        return super().load_state_dict(*a, **kw)

    def _apply(self, fn, *a, **kw):
        self._weights_version += 1
        return super()._apply(fn, *a, **kw)

    def mark_weights_changed(self):
        """Call after mutating parameters in place while staying in eval mode."""
        self._weights_version += 1

    # -- dispatch
    def _use_cuda_path(self, imgs) -> bool:
        return (not self.training) and (not torch.is_grad_enabled())

    def forward(self, imgs, proj_matrices, depth_values, filename=None):
        if self._use_cuda_path(imgs):
            from .engine import InferenceEngine  # imports the C ABI; raises if the library is missing
            if not imgs[0].is_cuda:
                from ._lib import MvsterLibraryError
                raise MvsterLibraryError("MVS4net inference runs on the sm_100a CUDA library only: move the model and "
                                         "inputs to a CUDA device (there is no CPU fallback)")
            dev = imgs[0].device
            eng = self._engines.get(dev.index)
            if eng is None:
                eng = self._engines[dev.index] = InferenceEngine(dev)
            if eng.weights_version != self._weights_version:
                eng.refresh_weights(self)
                eng.weights_version = self._weights_version
            if self.use_cuda_graph and (self._view_shard is None or os.environ.get("MVSTER_SHARD_GRAPH", "1") == "1"):
                return eng.forward_graphed(self, imgs, proj_matrices, depth_values, shard=self._view_shard)
            return eng.forward(self, imgs, proj_matrices, depth_values, shard=self._view_shard)
        return self._forward_autograd(imgs, proj_matrices, depth_values, filename)

    def _forward_autograd(self, imgs, proj_matrices, depth_values, filename=None):
        """MVS4Net.py:60-111 with PyTorch ops (training)."""
        depth_interval = (depth_values[:, -1] - depth_values[:, 0]) / depth_values.size(1)
        features = [self.feature(imgs[v]) for v in range(len(imgs))]
        outputs: Dict = {}
        prev = None
        for k in range(self.num_stage):
            key = f"stage{k + 1}"
            feats = [f[key] for f in features]
            B, C, H, W = feats[0].shape
            D = self.stage_splits[k]
            if k == 0:
                hypo = tp.inverse_range_init(depth_values, D, H, W) if self.inverse_depth else tp.linear_range_init(depth_values, D, H, W)
            elif self.inverse_depth:
                hypo = tp.inverse_range_schedule(prev["inverse_min_depth"].detach(), prev["inverse_max_depth"].detach(), D, H, W)
            else:
                hypo = tp.linear_range_schedule(prev["depth"].detach(), D, self.depth_interals_ratio[k] * depth_interval, H, W)
            prev = self.stagenet(feats, proj_matrices[key], depth_hypo=hypo, regnet=self.reg[k], stage_idx=k,
                                 group_cor=self.group_cor, group_cor_dim=self.group_cor_dim[k],
                                 split_itv=self.depth_interals_ratio[k], fn=filename)
            outputs[key] = prev
            outputs.update(prev)
        if self.mono and self.training:
            outputs = self.mono_depth_decoder(outputs, depth_values[:, 0], depth_values[:, 1])
        return outputs
