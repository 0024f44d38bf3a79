"""SIMT kernel SOURCES of libmvster_b200 executed on the CPU (tests/emu/simt_emu.h: one OS thread per CUDA thread, real
barriers for __syncthreads / warp votes / shuffles) and checked against the oracle - kernel logic coverage that needs no GPU:
  * et_fuse_win_kernel (csrc/et_fuse_win.cuh) for its three specialisations, window path and per-hypothesis fallback,
    partial / accumulate modes, ragged tiles;
  * geo_consistency_kernel (csrc/fusion.cu) against the reference's own outputs.
The GPU build of the same sources is byte-identical with or without this harness (only #ifdef MVSTER_CPU_EMU blocks were added)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

from util import GOLDEN, REPO, narrow_et_inputs, oracle
from oracle import fusion_oracle

EMU = REPO / "tests" / "emu"


@pytest.fixture(scope="module")
def emu():
    out = EMU / "_build"
    out.mkdir(exist_ok=True)
    lib = out / "libmvster_emu.so"
    srcs = [EMU / "kernels_emu.cpp", EMU / "simt_emu.h", REPO / "mvster_b200/csrc/et_fuse_win.cuh", REPO / "mvster_b200/csrc/et_fuse_tiled.cuh",
            REPO / "mvster_b200/csrc/et_args.cuh", REPO / "mvster_b200/csrc/fusion.cu"]
    if not lib.exists() or lib.stat().st_mtime < max(s.stat().st_mtime for s in srcs):
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-pthread", "-shared", "-fPIC", str(EMU / "kernels_emu.cpp"), "-o", str(lib)])
    return C.CDLL(str(lib))


def fptr(a):
    return a.ctypes.data_as(C.c_void_p)


def run_win(emu, feats, cams, hypo, G, flags=0, cost=None, wsum=None):
    """feats [ref, src...] each [B,C,H,W] torch; returns cost [B,G,D,H,W] numpy (and wsum)."""
    B, Cc, H, W = feats[0].shape
    D = hypo.shape[1]
    nhwc = [np.ascontiguousarray(f.permute(0, 2, 3, 1).numpy()) for f in feats]
    ref_full = oracle.compose_projection(cams[:, 0])
    pose = np.zeros((B, len(feats) - 1, 12), np.float32)
    for v in range(1, len(feats)):
        R, t = oracle.relative_pose(oracle.compose_projection(cams[:, v]), ref_full)
        pose[:, v - 1, :9] = R.reshape(B, 9).numpy()
        pose[:, v - 1, 9:] = t.reshape(B, 3).numpy()
    hy = np.ascontiguousarray(hypo.numpy())
    if cost is None:
        cost = np.full((B, D, H, W, G), np.nan, np.float32)
    if wsum is None:
        wsum = np.zeros((B, D, H, W), np.float32)
    srcs = (C.c_void_p * (len(feats) - 1))(*[a.ctypes.data for a in nhwc[1:]])
    rc = emu.emu_et_fuse_win(fptr(nhwc[0]), srcs, len(feats) - 1, fptr(pose), fptr(hy), fptr(cost), fptr(wsum),
                             B, Cc, G, D, H, W, H, W, C.c_float(2.0), flags)
    assert rc == 0
    return cost, wsum


EMU_CASES = [  # (B, nv, C, G, D, H, W, step_deg, rel_span): small grids (every CUDA thread is an OS thread here)
    (1, 3, 8, 4, 4, 8, 64, 1.0, 0.3),
    (1, 3, 8, 4, 4, 6, 40, 3.0, 0.12),    # ragged tile, wide baseline: out-of-image taps -> per-hypothesis path
    (1, 3, 16, 4, 4, 8, 32, 2.0, 0.2),    # two lanes per pixel (shuffles)
    (2, 2, 32, 8, 8, 4, 16, 4.0, 0.06),   # four lanes per pixel, D = 8, batch 2
]


@pytest.mark.parametrize("case", EMU_CASES)
def test_window_kernel_source_on_cpu_matches_oracle(emu, case):
    B, nv, Cc, G, D, H, W, step, span = case
    feats, cams, hypo = narrow_et_inputs(B, nv, Cc, D, H, W, step, span, seed=13)
    want = oracle.et_aggregate(feats, cams, hypo, True, G, 2.0).numpy()           # [B,G,D,H,W]
    cost, _ = run_win(emu, feats, cams, hypo, G)
    got = cost.transpose(0, 4, 1, 2, 3)
    assert np.isfinite(got).all()
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 2e-4 * scale


def test_window_kernel_source_partial_and_accumulate_on_cpu(emu):
    B, nv, Cc, G, D, H, W = 1, 4, 8, 4, 4, 8, 32
    feats, cams, hypo = narrow_et_inputs(B, nv, Cc, D, H, W, 2.0, 0.2, seed=17)
    acc_want, w_want = oracle.et_aggregate(feats, cams, hypo, True, G, 2.0, partial=True)
    c1, w1 = run_win(emu, feats[:3], cams[:, :3], hypo, G, flags=1)                                    # views 1-2, partial
    c1, w1 = run_win(emu, [feats[0], feats[3]], cams[:, [0, 3]], hypo, G, flags=1 | 2, cost=c1, wsum=w1)  # + view 3, accumulate
    assert np.abs(w1 - w_want.numpy()).max() < 1e-5
    got = c1.transpose(0, 4, 1, 2, 3)
    assert np.abs(got - acc_want.numpy()).max() <= 2e-4 * np.abs(acc_want.numpy()).max()


@pytest.mark.parametrize("name", ["plane_4v_48x64", "plane_3v_40x56_wide"])
def test_geo_consistency_kernel_source_on_cpu_matches_reference(emu, name):
    from mvster_b200.fusion import _prepare_mats
    z = np.load(GOLDEN / "fusion" / f"{name}.npz")
    nv, H, W, seed = (int(x) for x in z["meta"])
    views = fusion_oracle.synthetic_scene(nv, H, W, seed, float(z["step"]))
    ref = views[0]
    count, dsum = np.zeros((H, W), np.int32), np.zeros((H, W), np.float32)
    for v in range(1, nv):
        s = views[v]
        mats = np.ascontiguousarray(_prepare_mats(ref["K"], ref["E"], s["K"], s["E"]))
        dr, xs, ys = (np.empty((H, W), np.float32) for _ in range(3))
        mask = np.empty((H, W), np.uint8)
        rc = emu.emu_geo_consistency(fptr(ref["depth"]), fptr(s["depth"]), mats.ctypes.data_as(C.c_void_p), fptr(dr), fptr(xs), fptr(ys),
                                     fptr(mask), fptr(count), fptr(dsum), H, W, H, W, C.c_float(1.0), C.c_float(0.01))
        assert rc == 0
        assert np.array_equal(mask.astype(bool), z[f"mask{v}"])
        assert np.array_equal(dr, z[f"depth_reprojected{v}"])
        assert np.array_equal(xs, z[f"x2d_src{v}"], equal_nan=True) and np.array_equal(ys, z[f"y2d_src{v}"], equal_nan=True)
    assert np.array_equal(count, z["geo_mask_sum"])
    assert np.array_equal((dsum + ref["depth"]) / (count + 1), z["depth_est_averaged"])
