"""Shared helpers for the test-suite (the only place besides bench/smoke allowed to touch oracle/)."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
GOLDEN = REPO / "tests" / "golden"
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

from mvster_b200 import MVS4net, synth  # noqa: E402
from oracle import mvster_oracle as oracle  # noqa: E402

SHIPPED = dict(reg_net="reg2d", group_cor=True, group_cor_dim=[8, 8, 4, 4], inverse_depth=True, mono=True, attn_temp=2)

GOLDEN_CASES = {
    "shipped_b1_v3_64x128": SHIPPED,
    "shipped_b2_v2_64x64": SHIPPED,
    "reg3d_b1_v2_64x64": dict(reg_net="reg3d", group_cor=True, group_cor_dim=[8, 8, 4, 4], inverse_depth=True, attn_temp=2),
    "plain_b1_v2_64x64": dict(reg_net="reg2d", group_cor=False, inverse_depth=False, attn_fuse_d=False, attn_temp=2),
}


def oracle_cfg(kwargs: dict) -> dict:
    cfg = dict(oracle.DEFAULT_CFG)
    cfg.update(reg_net=kwargs.get("reg_net", "reg2d"), group_cor=kwargs.get("group_cor", False),
               group_cor_dim=kwargs.get("group_cor_dim", [8, 8, 8, 8]), inverse_depth=kwargs.get("inverse_depth", False),
               attn_temp=float(kwargs.get("attn_temp", 2)), attn_fuse_d=kwargs.get("attn_fuse_d", True),
               mono=kwargs.get("mono", False))
    return cfg


def build_model(kwargs: dict, seed: int) -> MVS4net:
    """Our module with the deterministic synthetic weights of mvster_b200.synth (same as the golden script)."""
    torch.manual_seed(0)
    m = MVS4net(**kwargs)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(synth.synthetic_state(shapes, seed), strict=True)
    return m.eval()


def load_golden(name: str):
    z = np.load(GOLDEN / f"{name}.npz")
    nv = int(z["meta_Nv"])
    imgs = [torch.from_numpy(z[f"img{v}"]) for v in range(nv)]
    proj = {f"stage{s}": torch.from_numpy(z[f"proj_stage{s}"]) for s in range(1, 5)}
    dv = torch.from_numpy(z["depth_values"])
    return z, imgs, proj, dv


def top2_gap(attn: torch.Tensor) -> torch.Tensor:
    """Difference between the two largest probabilities over D: [B,D,H,W] -> [B,H,W]."""
    t = attn.topk(2, dim=1).values
    return t[:, 0] - t[:, 1]


def narrow_et_inputs(B, nv, C, D, H, W, step_deg, rel_span, seed=0):
    """Inputs shaped like cascade stages 2-4: every pixel has its own hypothesis centre (425..935) and its D
    hypotheses cover only ``rel_span`` of it (far -> near), so they sample within one or two source-pixel
    cells - the regime the window kernel (csrc/et_fuse_win.cuh) is built for.  Returns (features, cams, hypo)."""
    rng = np.random.RandomState(seed)
    feats = [torch.from_numpy(rng.randn(B, C, H, W).astype(np.float32)) for _ in range(nv)]
    cams = synth.stage_projections(synth.arc_cameras(nv, H, W, step_deg), B, num_stage=1)["stage1"]
    centre = rng.uniform(450.0, 900.0, (B, 1, H, W))
    lin = np.linspace(0.5, -0.5, D).reshape(1, D, 1, 1)
    hypo = torch.from_numpy((centre * (1.0 + rel_span * lin)).astype(np.float32))
    return feats, cams, hypo


def warp_edge_through_et(capi, device="cpu"):
    """The reference's own ``homo_warping`` outputs on its edge-case fixture (tests/golden/warp_edge.npz: samples out of bounds on
    every side, a hypothesis with z exactly 0, source size != reference size) recovered THROUGH the fused warp + ET kernel: with one
    source view, C = G = 4 (one channel per group), an all-ones reference feature and a huge softmax temperature the cost volume
    is a / (1e-8 + a) * warped with a = 1/(4 * sqrt(4)), i.e. the warped volume to 8e-8 relative.  D = 3 in the fixture; the last
    hypothesis is duplicated to reach D = 4.  Returns (got [B,C,3,Hr,Wr], want)."""
    import numpy as np
    z = np.load(GOLDEN / "warp_edge.npz")
    src, hypo = torch.from_numpy(z["src"]), torch.from_numpy(z["hypo"])
    assert np.array_equal(z["ref_proj"], np.broadcast_to(np.eye(4, dtype=np.float32), z["ref_proj"].shape))
    P = torch.from_numpy(z["src_proj"])                                  # ref_proj = I  =>  src_proj @ inv(ref_proj) = src_proj
    pose = torch.cat([P[:, :3, :3].reshape(-1, 9), P[:, :3, 3]], 1)[:, None].contiguous()
    B, C, Hs, Ws = src.shape
    _, D, Hr, Wr = hypo.shape
    hypo4 = torch.cat([hypo, hypo[:, -1:]], 1).contiguous()
    ref = torch.ones(B, Hr, Wr, C)
    outs = []
    for kw in (dict(generic=True), dict(window=False), dict(window=True)):
        cost = capi.et_fuse(ref.to(device), [src.permute(0, 2, 3, 1).contiguous().to(device)], pose.to(device), hypo4.to(device),
                            C, 1e6, **kw)                               # [B,4,Hr,Wr,C]
        outs.append(cost.permute(0, 4, 1, 2, 3)[:, :, :D].cpu())
    return outs, torch.from_numpy(z["warped"])
