"""Timed A/B of the storage modes of the whole forward (CUDA events, L2 flushed between steps, module defaults = CUDA-graph replay):
fp32 storage | bf16 storage with fp32 containers between the regulariser's layers | bf16 storage with packed operands.
    python tools/bf16_ab.py [H W views steps]"""
import json
import os
import statistics
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from mvster_b200 import synth  # noqa: E402


def main():
    H, W, nv, steps = (int(a) for a in (sys.argv[1:5] + ["1152", "1600", "5", "8"][len(sys.argv) - 1:]))
    dev = torch.device("cuda", 0)
    model = bench.build_model(dev)
    flush = torch.empty(128 << 20, device=dev)
    imgs, proj, dv = synth.make_inputs(1, nv, H, W, seed=0)
    imgs, proj, dv = [t.to(dev) for t in imgs], {k: v.to(dev) for k, v in proj.items()}, dv.to(dev)
    res = {"H": H, "W": W, "views": nv}
    outs = {}
    for name, storage, packed in (("fp32", "fp32", "1"), ("bf16_container", "bf16", "0"), ("bf16_packed", "bf16", "1")):
        model.storage = storage
        os.environ["MVSTER_BF16_PACKED"] = packed
        eng = model._engines.get(dev.index)
        if eng is not None:
            eng._graphs.clear()

        def fwd():
            with torch.no_grad():
                return model(imgs, proj, dv)
        ms, best = bench._time_forward(fwd, steps, flush)
        outs[name] = {k: v.clone() for k, v in fwd()["stage4"].items() if k in ("depth", "attn_weight")}
        res[name] = {"ms_per_step": round(ms, 4), "min_ms": round(best, 4)}
    a, b = outs["bf16_container"], outs["bf16_packed"]
    res["packed_equals_container"] = bool(torch.equal(a["depth"], b["depth"]) and torch.equal(a["attn_weight"], b["attn_weight"]))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
