"""tcgen05 (tensor-core) convolution vs the exact-fp32 CUDA-core convolution, one subprocess per case."""
import json
import subprocess
import sys

import pytest

from util import REPO

pytestmark = pytest.mark.gpu

CASES = [  # cin cout kd B D H W npass [flags]
    "32 32 3 1 4 16 32 3",
    "16 16 3 1 8 24 40 3 skip",      # 64-byte swizzle path, ragged right/bottom tiles
    "64 64 3 1 8 8 10 3",            # two K chunks per tap, image smaller than one tile
    "32 32 1 2 4 16 16 3 norelu",    # (1,3,3) kernel, batch 2
    "32 64 3 1 4 32 48 1",           # plain TF32
    "16 16 3 1 4 256 320 3",         # 2560 tiles: 4 accumulators per CTA in TMEM
    "64 64 3 1 4 64 80 3 skip",      # reg2d conv6 at cfg2 stage 4
    "reg2d 8 1 8 64 80 3",           # whole reg2d U-Net, stage-1 shape of cfg2, 3xTF32
    "reg2d 4 1 4 512 640 3",         # stage-4 shape of cfg2 (1.31 M voxels)
    "reg2d 4 1 4 128 160 1",         # plain TF32
]


@pytest.mark.parametrize("case", CASES)
def test_tc_conv_matches_exact_conv(case):
    p = subprocess.run([sys.executable, str(REPO / "tests" / "tc_conv_check.py"), *case.split()], capture_output=True, text=True,
                       timeout=180)
    assert p.returncode == 0, f"subprocess failed:\n{p.stdout[-2000:]}\n{p.stderr[-3000:]}"
    res = json.loads(p.stdout.strip().splitlines()[-1])
    out = REPO / "gpurun_out"
    out.mkdir(exist_ok=True)
    with open(out / "tc_conv_report.jsonl", "a") as f:
        f.write(json.dumps(res) + "\n")
    assert res["finite"]
    npass = int(case.split()[-1]) if case.startswith("reg2d") else int(case.split()[7])
    tol = (2e-5 if case.startswith("reg2d") else 3e-6) if npass == 3 else 5e-3
    assert res["rel"] < tol, res
