"""tcgen05 (tensor-core) convolutions vs the exact-fp32 CUDA-core convolution, one subprocess per case.

Tolerances (max abs error / max |output|): 3xTF32 3e-5 - the per-product error of the hi/lo split is 2^-21,
the tensor core accumulates K = 27*Cin products in fp32 with truncation; measured 3e-6 .. 1.6e-5.  Plain TF32 5e-3."""
import json
import subprocess
import sys

import pytest

from util import REPO

pytestmark = pytest.mark.gpu

CASES = [  # generation cin cout kd B D H W npass [flags]
    "v2 32 32 3 1 4 16 32 3",
    "v2 16 16 3 1 8 24 40 3 skip",      # ragged right/bottom tiles (16x8 tiles on 24x40)
    "v2 64 64 3 1 8 8 10 3",            # four 16-channel chunks, image smaller than one tile, depth padding
    "v2 32 32 1 2 4 16 16 3 norelu",    # (1,3,3) kernel, batch 2
    "v2 32 64 3 1 4 32 48 1",           # plain TF32
    "v2 64 8 1 2 1 32 40 3 norelu",     # 2-D 3x3 conv 64 -> 8 (N padded to 16): the FPN out4 shape class
    "v2 16 16 3 1 4 256 320 3",         # reg2d conv2 at cfg2 stage 4: 2560 tiles, 4 accumulators per CTA
    "v2 32 32 3 1 4 128 160 3 skip",    # conv4
    "v2 64 64 3 1 4 64 80 3 skip",      # conv6
    # generation 3 (persistent, 3 x bf16, fp32-faithful): v3 cin cout kd k stride B D H W [flags]
    "v3 16 16 1 3 1 1 2 24 40 skip",    # ragged tiles, 2 planes
    "v3 16 16 3 3 1 1 8 24 40",         # depth taps with plane skipping at the volume border
    "v3 32 32 3 3 1 2 4 16 16 norelu",  # two channel chunks, batch 2
    "v3 64 64 3 3 1 1 4 64 80 skip",    # conv6 at cfg2 stage 4 (N = 64: one tile per group)
    "v3 64 8 1 3 1 2 1 32 40 norelu",   # Cout = 8 padded to N = 16
    "v3 8 8 1 3 1 1 4 32 48",           # Cin = 8: half-filled K (FPN conv0.1, reg2d conv0 at G = 8)
    "v3 4 8 1 3 1 1 4 32 48",           # Cin = 4 (reg2d conv0 at G = 4)
    "v3 32 64 1 1 1 1 1 64 80 norelu",  # 1x1 = plain GEMM
    "v3 8 16 1 3 2 1 4 64 80",          # stride (1,2,2) 3x3: four parity classes (reg2d conv1)
    "v3 16 32 1 3 2 1 4 30 44",         # ... odd output size (15 x 22), ragged
    "v3 32 64 1 3 2 2 2 32 32",         # conv5
    "v3 8 16 1 5 2 2 1 64 96",          # FPN conv1.0: 5x5 stride 2
    "v3 16 32 1 5 2 1 1 50 70",         # FPN conv2.0, ragged
    "v3 32 64 1 5 2 1 1 64 80",         # FPN conv3.0
    "v3 16 16 3 3 1 1 4 256 320",       # reg2d conv2 at cfg2 stage 4: 2560 tiles, persistent loop
    "v3 16 16 1 3 1 5 1 512 640",       # FPN out3 (12800 tiles)
    "v3 32 32 3 3 1 1 4 128 160 skip",  # conv4
    # generation-3 transposed conv (1,3,3)/stride (1,2,2) = 2x2 conv + depth-to-space: d3 cin cout B D H W [skip]
    "d3 16 8 1 2 24 40 skip",           # conv11 class: four parity classes in one launch (N = 32), ragged tiles
    "d3 32 16 2 2 16 16",               # conv9 class (N = 64), batch 2
    "d3 64 32 1 4 8 10 skip",           # conv7 class: two launches (output rows of parity 0, 1), image smaller than a tile
    "d3 16 8 1 4 256 320 skip",         # conv11 at cfg2 stage 4
    "d3 32 16 1 4 128 160 skip",        # conv9
    "d3 64 32 1 4 64 80 skip",          # conv7
    # the same kernel with two fp16 terms per operand (MVSTER_TC3_FP16X2): same fp64 yardstick, same tolerance
    "v3 16 16 3 3 1 1 8 24 40 h16",
    "v3 32 32 3 3 1 2 4 16 16 norelu h16",
    "v3 64 64 3 3 1 1 4 64 80 skip h16",
    "v3 64 8 1 3 1 2 1 32 40 norelu h16",
    "v3 8 8 1 3 1 1 4 32 48 h16",
    "v3 4 8 1 3 1 1 4 32 48 h16",
    "v3 32 64 1 1 1 1 1 64 80 norelu h16",
    "v3 16 32 1 3 2 1 4 30 44 h16",
    "v3 8 16 1 5 2 2 1 64 96 h16",
    "v3 32 64 1 5 2 1 1 64 80 h16",
    "v3 16 16 3 3 1 1 4 256 320 h16",
    "v3 16 16 1 3 1 5 1 512 640 h16",
    "v3 32 32 3 3 1 1 4 128 160 skip h16",
    "d3 16 8 1 2 24 40 skip h16",
    "d3 32 16 2 2 16 16 h16",
    "d3 64 32 1 4 8 10 skip h16",
    "d3 16 8 1 4 256 320 skip h16",
    # packed fp16-pair operands (mvster_conv_tc3_pb16 with MVSTER_TC3_FP16X2: x and skip as (a1, a2) planes in the operand's own
    # order, TMA straight into the operand ring, no converter warps); p16f = fp32 output, p16 = packed output (split by the epilogue)
    "v3 16 16 3 3 1 1 8 24 40 skip h16 p16f",   # depth taps, ragged tiles, 4-D merged map
    "v3 8 16 1 3 2 1 4 64 80 h16 p16f",         # conv1: stride 2 (5-D map with element strides), one octet, two taps per MMA
    "v3 16 32 1 3 2 1 4 30 44 h16 p16",         # conv3, odd output size
    "v3 32 64 1 3 2 2 2 32 32 h16 p16",         # conv5
    "v3 32 32 3 3 1 2 4 16 16 norelu h16 p16",  # two channel chunks, batch 2
    "v3 64 64 3 3 1 1 4 64 80 skip h16 p16",    # conv6
    "v3 16 16 3 3 1 1 4 256 320 h16 p16",       # conv2 at cfg2 stage 4
    "v3 32 32 3 3 1 1 4 128 160 skip h16 p16",  # conv4 (streamed weights)
    "d3 16 8 1 2 24 40 skip h16 p16f",          # conv11: packed skip, fp32 out
    "d3 32 16 2 2 16 16 skip h16 p16",          # conv9
    "d3 64 32 1 4 8 10 skip h16 p16",           # conv7: two launches into one packed output
    "d3 16 8 1 4 256 320 skip h16 p16f",        # conv11 at cfg2 stage 4
    "reg2dv2 8 1 8 64 80 3",            # whole reg2d U-Net, stage-1 shape of cfg2, 3xTF32
    "reg2dv2 4 1 4 512 640 3",          # stage-4 shape of cfg2 (1.31 M voxels)
    "reg2dv2 4 1 4 128 160 1",          # plain TF32
]


@pytest.mark.parametrize("case", CASES)
def test_tc_conv_matches_exact_conv(case):
    p = subprocess.run([sys.executable, str(REPO / "tests" / "tc_conv_check.py"), *case.split()], capture_output=True, text=True,
                       timeout=90)
    assert p.returncode == 0, f"subprocess failed:\n{p.stdout[-2000:]}\n{p.stderr[-3000:]}"
    res = json.loads(p.stdout.strip().splitlines()[-1])
    out = REPO / "gpurun_out"
    out.mkdir(exist_ok=True)
    with open(out / "tc_conv_report.jsonl", "a") as f:
        f.write(json.dumps(res) + "\n")
    assert res["finite"]
    f = case.split()
    if f[0] in ("v3", "d3"):  # measured against an fp64 convolution: fp32-level agreement
        assert res["rel"] < 1e-5, res  # measured 3e-7 .. 4.5e-6 (K up to 1728 products, fp32 accumulation in the tensor core)
        return
    npass = int(f[6]) if f[0].startswith("reg2d") else int(f[8])
    tol = (1e-4 if f[0].startswith("reg2d") else 3e-5) if npass == 3 else 5e-3
    assert res["rel"] < tol, res
