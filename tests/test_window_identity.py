"""CPU check of the identity behind the window kernel (mvster_b200/csrc/et_fuse_win.cuh): correlating the
reference pixel with every tap of a 3 x 3 source block first and interpolating the per-group correlations
piecewise-linearly afterwards equals the reference's warp-then-correlate (mvs4net_utils.py:13-59, :1037-1040)
whenever the D hypotheses of the pixel stay within two cells per axis, straddling lanes included."""
import numpy as np
import pytest
import torch

from util import narrow_et_inputs, oracle


def window_group_correlation(ref, src, R, t, hypo, G):
    """Torch restatement of the kernel's fast path for one source view.  ref/src [B,C,H,W]; R [B,3,3], t [B,3,1];
    hypo [B,D,H,W] -> (cor [B,G,D,H,W], fits [B,H,W]); cor is only meaningful where fits."""
    B, C, H, W = ref.shape
    D = hypo.shape[1]
    Hs, Ws = src.shape[2:]
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    pix = torch.stack((xx, yy, torch.ones_like(xx)), 0).reshape(1, 1, 3, H, W)
    ray = (R.reshape(B, 3, 3, 1, 1) * pix).sum(2)                         # [B,3,H,W]
    pts = ray.unsqueeze(2) * hypo.unsqueeze(1) + t.reshape(B, 3, 1, 1, 1)  # [B,3,D,H,W]
    z = pts[:, 2]
    z = torch.where(z == 0, torch.full_like(z, 1e-9), z)
    ix, iy = pts[:, 0] / z, pts[:, 1] / z                                  # [B,D,H,W]
    fx, fy = ix.floor(), iy.floor()
    bx, mx = fx.min(1).values, fx.max(1).values
    by, my = fy.min(1).values, fy.max(1).values
    fits = (bx >= 0) & (mx < Ws - 1) & (by >= 0) & (my < Hs - 1) & (mx - bx <= 1) & (my - by <= 1)
    bxi, byi = bx.clamp(0, Ws - 1).long(), by.clamp(0, Hs - 1).long()  # only used where fits
    refg = ref.reshape(B, G, C // G, H, W) / (C // G)
    T = torch.zeros(3, 3, B, G, H, W)
    bidx = torch.arange(B).reshape(B, 1, 1).expand(B, H, W)
    for r in range(3):
        for c in range(3):
            tap = src[bidx, :, (byi + r).clamp(max=Hs - 1), (bxi + c).clamp(max=Ws - 1)].permute(0, 3, 1, 2)  # [B,C,H,W]
            T[r, c] = (refg * tap.reshape(B, G, C // G, H, W)).sum(2)
    # lanes that do not straddle never load the third column / row: the kernel holds zeros there
    needx, needy = (mx > bx).reshape(B, 1, H, W), (my > by).reshape(B, 1, H, W)
    T[:, 2] = T[:, 2] * needx
    T[2, :] = T[2, :] * needy
    ux, uy = ix - bx.unsqueeze(1), iy - by.unsqueeze(1)                    # [B,D,H,W]
    ax, bxw = ux.clamp(max=1.0), (ux - 1.0).clamp(min=0.0)
    ay, byw = uy.clamp(max=1.0), (uy - 1.0).clamp(min=0.0)
    h = [T[r, 0].unsqueeze(2) + ax.unsqueeze(1) * (T[r, 1] - T[r, 0]).unsqueeze(2)
         + bxw.unsqueeze(1) * (T[r, 2] - T[r, 1]).unsqueeze(2) for r in range(3)]   # 3 x [B,G,D,H,W]
    cor = h[0] + ay.unsqueeze(1) * (h[1] - h[0]) + byw.unsqueeze(1) * (h[2] - h[1])
    return cor, fits


@pytest.mark.parametrize("C,G,D,rel_span,step", [(8, 4, 4, 0.15, 1.0), (16, 4, 4, 0.2, 2.0), (32, 8, 8, 0.25, 1.0)])
def test_correlate_then_interpolate_equals_warp_then_correlate(C, G, D, rel_span, step):
    B, nv, H, W = 1, 3, 24, 40
    feats, cams, hypo = narrow_et_inputs(B, nv, C, D, H, W, step, rel_span, seed=5)
    ref_full = oracle.compose_projection(cams[:, 0])
    covered = 0
    for v in range(1, nv):
        src_full = oracle.compose_projection(cams[:, v])
        warped = oracle.plane_sweep_warp(feats[v], src_full, ref_full, hypo)
        want = (warped.reshape(B, G, C // G, D, H, W) * feats[0].reshape(B, G, C // G, 1, H, W)).mean(2)
        R, t = oracle.relative_pose(src_full, ref_full)
        got, fits = window_group_correlation(feats[0], feats[v], R, t, hypo, G)
        m = fits.reshape(B, 1, 1, H, W).expand_as(want)
        covered += int(fits.sum())
        err = (got - want)[m].abs().max().item()
        assert err < 2e-4 * want.abs().max().item(), err
    assert covered > 0.5 * B * H * W * (nv - 1)
