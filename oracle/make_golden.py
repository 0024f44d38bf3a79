"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference, present only in the build container) on seeded synthetic inputs.

    python oracle/make_golden.py            # rewrites every fixture

The fixtures pin oracle/mvster_oracle.py (tests/test_oracle_golden.py) and the loss
mirror in mvster_b200/losses.py.  Weights are NOT stored: they are regenerated from
mvster_b200.synth.synthetic_state(shapes, seed), which is numpy-RandomState deterministic.
This script is test infrastructure; nothing at run time on the GPU box reads /root/reference.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
REF = Path(os.environ.get("MVSTER_REFERENCE", "/root/reference"))
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REF))

import models as ref_models  # noqa: E402  (the reference package)
from models import mvs4net_utils as ref_utils  # noqa: E402
from mvster_b200 import synth  # noqa: E402

OUT = REPO / "tests" / "golden"

CASES = {
    # name: (ctor kwargs, B, Nv, H, W, seed, step_deg)
    "shipped_b1_v3_64x128": (dict(reg_net="reg2d", group_cor=True, group_cor_dim=[8, 8, 4, 4], inverse_depth=True,
                                  mono=True, attn_temp=2), 1, 3, 64, 128, 11, 1.0),
    "shipped_b2_v2_64x64": (dict(reg_net="reg2d", group_cor=True, group_cor_dim=[8, 8, 4, 4], inverse_depth=True,
                                 mono=True, attn_temp=2), 2, 2, 64, 64, 12, 5.0),
    "reg3d_b1_v2_64x64": (dict(reg_net="reg3d", group_cor=True, group_cor_dim=[8, 8, 4, 4], inverse_depth=True,
                               attn_temp=2), 1, 2, 64, 64, 13, 2.0),
    "plain_b1_v2_64x64": (dict(reg_net="reg2d", group_cor=False, inverse_depth=False, attn_fuse_d=False,
                               attn_temp=2), 1, 2, 64, 64, 14, 2.0),
}


def build_reference(kwargs, seed):
    torch.manual_seed(0)
    model = ref_models.MVS4net(**kwargs)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(synth.synthetic_state(shapes, seed), strict=True)
    return model.eval()


def run_case(name, spec):
    kwargs, B, Nv, H, W, seed, step = spec
    model = build_reference(kwargs, seed)
    imgs, proj, dv = synth.make_inputs(B, Nv, H, W, seed=seed, step_deg=step)
    costs = {}
    hooks = [model.reg[k].register_forward_pre_hook(lambda m, inp, k=k: costs.__setitem__(k, inp[0].detach().clone()))
             for k in range(4)]
    with torch.no_grad():
        out = model(imgs, proj, dv)
        feats0 = model.feature(imgs[0])
    for h in hooks:
        h.remove()
    blob = {"meta_B": B, "meta_Nv": Nv, "meta_H": H, "meta_W": W, "meta_seed": seed, "meta_step": step}
    for v, im in enumerate(imgs):
        blob[f"img{v}"] = im.numpy()
    for k, p in proj.items():
        blob[f"proj_{k}"] = p.numpy()
    blob["depth_values"] = dv.numpy()
    for s in range(1, 5):
        st = out[f"stage{s}"]
        for key in ("depth", "photometric_confidence", "hypo_depth", "attn_weight", "inverse_min_depth", "inverse_max_depth"):
            if key in st:
                blob[f"s{s}_{key}"] = st[key].numpy()
        blob[f"s{s}_cost"] = costs[s - 1].numpy()
        blob[f"fpn_view0_stage{s}"] = feats0[f"stage{s}"].numpy()
    np.savez_compressed(OUT / f"{name}.npz", **blob)
    return model, out, (imgs, proj, dv)


def warp_edge_case():
    """homo_warping called directly (mvs4net_utils.py:13-59) with poses that push samples out of
    bounds on every side, one hypothesis that makes z exactly 0, and non-square source size."""
    rng = np.random.RandomState(5)
    B, C, Hs, Ws, Hr, Wr, D = 2, 4, 9, 14, 7, 10, 3
    src = torch.from_numpy(rng.randn(B, C, Hs, Ws).astype(np.float32))
    ref_proj = torch.eye(4).repeat(B, 1, 1)
    src_proj = torch.eye(4).repeat(B, 1, 1)
    src_proj[0, :3, :3] = torch.tensor([[1.1, 0.2, -3.0], [-0.1, 0.9, 2.0], [0.0, 0.0, 1.0]])
    src_proj[0, :3, 3] = torch.tensor([4.0, -2.5, 0.0])
    src_proj[1, :3, :3] = torch.tensor([[0.7, 0.0, 1.0], [0.0, 1.3, -1.0], [0.01, 0.0, 1.0]])
    src_proj[1, :3, 3] = torch.tensor([-30.0, 8.0, -2.0])  # z = 0.01*x*d + d - 2 -> 0 at x=0, d=2
    hypo = torch.tensor([1.0, 2.0, 3.5]).view(1, D, 1, 1).repeat(B, 1, Hr, Wr).contiguous()
    hypo[0, 2] += torch.from_numpy(rng.rand(Hr, Wr).astype(np.float32))
    out = ref_utils.homo_warping(src, src_proj, ref_proj, hypo)
    np.savez_compressed(OUT / "warp_edge.npz", src=src.numpy(), src_proj=src_proj.numpy(), ref_proj=ref_proj.numpy(),
                        hypo=hypo.numpy(), warped=out.numpy())


def loss_case(model, out, inputs):
    """MVS4net_loss / Blend_loss / sinkhorn on the shipped-config outputs with synthetic ground truth."""
    rng = np.random.RandomState(21)
    gt, mask = {}, {}
    for s in range(1, 5):
        d = out[f"stage{s}"]["depth"]
        gt[f"stage{s}"] = (d + torch.from_numpy(rng.normal(0, 8.0, tuple(d.shape)).astype(np.float32))).clamp(430, 930)
        mask[f"stage{s}"] = torch.from_numpy((rng.rand(*d.shape) > 0.3).astype(np.float32))
    blob = {}
    for s in range(1, 5):
        blob[f"gt_stage{s}"] = gt[f"stage{s}"].numpy()
        blob[f"mask_stage{s}"] = mask[f"stage{s}"].numpy()
    stage_only = {k: v for k, v in out.items() if k.startswith("stage")}
    for tag, kw in (("disc", dict(inverse_depth=True, ot_iter=3, ot_eps=1, ot_continous=False)),
                    ("cont", dict(inverse_depth=True, ot_iter=5, ot_eps=0.5, ot_continous=True, stage_lw=[1, 2, 3, 4]))):
        total, l1, ot, oor = ref_models.MVS4net_loss(stage_only, gt, mask, **kw)
        blob[f"{tag}_total"] = np.float32(total.item())
        blob[f"{tag}_ot"] = np.array([x.item() for x in ot], np.float32)
        blob[f"{tag}_oor"] = np.array([x.item() for x in oor], np.float32)
    res = ref_models.Blend_loss(stage_only, gt, mask, inverse_depth=True, depth_max=torch.tensor([935.0]),
                                depth_min=torch.tensor([425.0]))
    blob["blend_total"] = np.float32(res[0].item())
    blob["blend_metrics"] = np.array([res[4].item(), res[5].item(), res[6].item()], np.float32)
    np.savez_compressed(OUT / "losses.npz", **blob)


def train_mode_case():
    """Training-mode forward (mono decoder active, confidence == 0) for the autograd path mirror."""
    kwargs, B, Nv, H, W, seed, step = CASES["shipped_b1_v3_64x128"]
    model = build_reference(kwargs, seed).train()
    for m in model.modules():  # keep BN in inference behaviour so the result is a pure function of the weights
        if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
            m.eval()
    imgs, proj, dv = synth.make_inputs(B, 2, H, W, seed=seed, step_deg=step)
    with torch.no_grad():
        out = model(imgs, proj, dv)
    blob = {}
    for s in range(1, 5):
        blob[f"s{s}_depth"] = out[f"stage{s}"]["depth"].numpy()
        blob[f"s{s}_attn_weight"] = out[f"stage{s}"]["attn_weight"].numpy()
        if "mono_depth" in out[f"stage{s}"]:
            blob[f"s{s}_mono_depth"] = out[f"stage{s}"]["mono_depth"].numpy()
    np.savez_compressed(OUT / "train_mode.npz", **blob)


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    first = None
    for name, spec in CASES.items():
        res = run_case(name, spec)
        if first is None:
            first = res
        print("wrote", name)
    warp_edge_case()
    loss_case(*first)
    train_mode_case()
    total = sum(p.stat().st_size for p in OUT.glob("*.npz"))
    print(f"golden fixtures: {total / 1e6:.2f} MB in {OUT}")


if __name__ == "__main__":
    main()
