#!/usr/bin/env python
"""Benchmark of the MVSTER forward hot path on B200 (see DESIGN.md "Measurement").

    python bench.py --gpus 1 --steps 20 --warmup 5            # our arm
    python bench.py --impl reference --steps 3 --warmup 1     # the reference algorithm on the host CPU (oracle port)
    torchrun ... bench.py --gpus N ...                        # N ranks, one per GPU (batch-sharded replicas, weak scaling)

A step = one full ``MVS4net.forward`` (FPN4 features for all views + the 4-stage cascade) on one
synthetic frame per GPU: 5 views, 512x640, fp32, shipped configuration.  ``value`` times it with
the images already resident in HBM; ``e2e`` times the same public call starting from pinned host
images and ending with depth + confidence back in pinned host memory.  One JSON line on stdout.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

SHIPPED = dict(reg_net="reg2d", group_cor=True, group_cor_dim=[8, 8, 4, 4], inverse_depth=True, mono=True, attn_temp=2)
WORKLOAD = dict(views=5, H=512, W=640, batch_per_gpu=1, stages=4)
C_K, D_K, G_K = (64, 32, 16, 8), (8, 8, 4, 4), (8, 8, 4, 4)
METRIC, UNIT = "depth-maps/sec fwd, 5-view 512x640 4-stage; warp HBM GB/s vs B200 peak", "depth-maps/s"


def et_algorithmic_bytes(k: int, B: int, V: int, H: int, W: int, partial: bool = False) -> int:
    """every feature element once + hypotheses once + cost volume once (fp32)."""
    h, w = H >> (3 - k), W >> (3 - k)
    b = (1 + V) * B * C_K[k] * h * w * 4 + B * D_K[k] * h * w * 4 + B * G_K[k] * D_K[k] * h * w * 4
    return b + (B * D_K[k] * h * w * 4 if partial else 0)


def measured_peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            j = json.loads(p.read_text())
            return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def tensor_peak():
    """dense bf16 tensor throughput (TFLOP/s): the burst figure of MEASURED_PEAKS.json (kernel timed alone), else the recipe's fallback."""
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
        except Exception:
            pass
    return 1590.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        busy = [s for s in sm if s > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def pick_cpu_threads(run_small) -> int:
    """The oracle is plain torch-CPU; on many-core hosts the default (all cores) can be far slower
    than a moderate thread count.  Time a small forward at a few counts and keep the fastest."""
    cores = os.cpu_count() or 1
    best, best_t = cores, None
    for n in sorted({c for c in (8, 16, 32, 64, cores) if c <= cores}):
        torch.set_num_threads(n)
        run_small()
        t0 = time.perf_counter()
        run_small()
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = n, dt
    torch.set_num_threads(best)
    return best


def build_model(device, seed=0):
    from mvster_b200 import MVS4net, synth
    torch.manual_seed(0)
    m = MVS4net(**SHIPPED)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(synth.synthetic_state(shapes, seed), strict=True)
    return m.eval().to(device)


def run_reference(args, rank: int):
    """The reference algorithm (CPU oracle port, validated against the unmodified reference by
    tests/test_oracle_golden.py) timed on the host cores with all threads."""
    if rank != 0:
        return
    from mvster_b200 import MVS4net, synth
    from oracle import mvster_oracle as oracle
    torch.manual_seed(0)
    m = MVS4net(**SHIPPED)
    sd = synth.synthetic_state({k: tuple(v.shape) for k, v in m.state_dict().items()}, 0)
    imgs, proj, dv = synth.make_inputs(WORKLOAD["batch_per_gpu"], WORKLOAD["views"], WORKLOAD["H"], WORKLOAD["W"], seed=0)
    cfg = dict(oracle.DEFAULT_CFG)
    small = synth.make_inputs(1, 3, 128, 192, seed=1)
    cores = pick_cpu_threads(lambda: oracle.cascade_forward(sd, cfg, *small))
    for _ in range(args.warmup):
        oracle.cascade_forward(sd, cfg, imgs, proj, dv)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.cascade_forward(sd, cfg, imgs, proj, dv)
    dt = (time.perf_counter() - t0) / args.steps
    val = WORKLOAD["batch_per_gpu"] / dt
    sample = f"{args.steps} full 5-view 512x640 forwards (FPN4 + 4 stages) after {args.warmup} warm-up, torch CPU fp32"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_dict(1),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def config_dict(n, view_parallel=1):
    par = f"batch-sharded replicas x{n}" if view_parallel == 1 else \
        f"{n // view_parallel} frame replica(s) x {view_parallel}-way source-view sharding (1 NCCL all-reduce per stage)"
    return {"workload": "cfg2: DTU-mid 5-view 512x640, 4-stage cascade, fp32, shipped config (reg2d, group_cor 8/8/4/4, "
                        "D 8/8/4/4, inverse depth); one frame per GPU (per view group with --view-parallel)",
            "views": 5, "H": 512, "W": 640, "global_batch": n // view_parallel, "parallelism": par,
            "l2": "512 MB buffer written between timed steps (L2 flush)"}


def _time_forward(fn, steps, flush, warmup=3):
    """ms per call of fn: CUDA events on the current stream, L2 flushed before every call."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(steps):
        flush.fill_(1.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    ts = [s.elapsed_time(e) for s, e in evs]
    return statistics.mean(ts), min(ts)


def gpu_eager_baseline(model, imgs_d, proj_d, dv_d, flush, steps=10):
    """The >= 10x target's denominator (SURVEY.md 8d "GPU baseline"): the reference's forward as PyTorch ops, eager on this GPU
    through cuDNN/ATen, `cudnn.benchmark=True` as `test_mvs4.py:20`, with TF32 allowed (PyTorch's cuDNN default = how the
    reference runs) and with strict fp32.  /root/reference does not travel to the GPU box, so the op sequence is this
    repository's differentiable restatement of it (mvster_b200/torch_path.py + the module's own nn.Conv/BatchNorm layers:
    same ATen calls per view and stage as mvs4net_utils.py:13-59, 1015-1094), pinned by tests/test_host_cpu.py."""
    res = {"formulation": "reference op sequence as PyTorch ops (network._forward_autograd, eval, no_grad), eager, cudnn.benchmark=True"}
    B = imgs_d[0].shape[0]
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32

        def fwd():
            with torch.no_grad():
                return model._forward_autograd(imgs_d, proj_d, dv_d)
        ms, best = _time_forward(fwd, steps, flush, warmup=6)
        res["tf32_on" if tf32 else "tf32_off"] = {"value": B / (ms * 1e-3), "ms_per_step": ms, "min_ms": best, "allow_tf32": tf32}
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return res


def extra_legs(model, dev, flush, steps=5):
    """Informational legs on the same line: the reference's own test-time sizes (README.md:77-81, test_mvs4.py:41-42: 832x1152
    and 1152x1600, 5 views) through the default (CUDA-graph) path, and cfg2 through eager launches (no CUDA graph)."""
    from mvster_b200 import synth
    out = {"sizes": []}
    for (nv, H, W) in ((5, 832, 1152), (5, 1152, 1600)):
        imgs, proj, dv = synth.make_inputs(1, nv, H, W, seed=0)
        imgs = [t.to(dev) for t in imgs]
        proj = {k: v.to(dev) for k, v in proj.items()}
        dv = dv.to(dev)

        def fwd():
            with torch.no_grad():
                return model(imgs, proj, dv)
        ms, best = _time_forward(fwd, steps, flush)
        out["sizes"].append({"views": nv, "H": H, "W": W, "ms_per_step": ms, "min_ms": best, "value": 1.0 / (ms * 1e-3),
                             "cuda_graph": bool(model.use_cuda_graph)})
        del imgs, proj, dv
        eng = model._engines.get(dev.index)
        if eng is not None:
            eng._graphs.clear()  # each captured graph pins its workspace
        torch.cuda.empty_cache()
    return out


def bf16_storage_leg(model, dev, flush, ref_fp32=None, inputs_cfg2=None, steps=5):
    """BASELINE configs[2]: the bf16-storage configuration (model.storage = "bf16": bf16 features and cost volume in HBM, one-term
    bf16 tcgen05 convolutions in the regulariser, fp32 geometry / softmax / accumulation; the reference has no bf16 path, so
    PARITY IS UNPINNED - the yardstick is this repository's oracle.cascade_forward(storage_dtype=bfloat16,
    storage_fpn_internal=False)) at the reference's full test-time size, 5 views 1152x1600 (general_eval4.py:92-109 crops DTU's
    1200x1600 to a multiple of 64), with the warp/ET kernel's own roofline (bf16 bytes) and, at cfg2's size, the comparison with
    the bf16 oracle."""
    from mvster_b200 import capi, fpn_engine, synth
    peak, peak_src = measured_peaks()
    nv, H, W = 5, 1152, 1600
    res = {"workload": f"cfg3: {nv} views, {H}x{W}, 4-stage cascade, bf16 storage (features, cost volume, regulariser operands), one frame",
           "parity_pin": "unpinned (the reference has no bf16 path; oracle definition: oracle/mvster_oracle.py storage)"}
    prev = getattr(model, "storage", "fp32")
    eng = model._engines.get(dev.index)
    try:
        for storage in ("bf16", "fp32"):
            model.storage = storage
            imgs, proj, dv = synth.make_inputs(1, nv, H, W, seed=0)
            imgs, proj, dv = [t.to(dev) for t in imgs], {k: v.to(dev) for k, v in proj.items()}, dv.to(dev)

            def fwd():
                with torch.no_grad():
                    return model(imgs, proj, dv)
            ms, best = _time_forward(fwd, steps, flush)
            res["ms_per_step" if storage == "bf16" else "fp32_ms_per_step"] = ms
            if storage == "bf16":
                res.update(min_ms=best, value=1.0 / (ms * 1e-3), unit=UNIT)
                out = fwd()
                # the stage launches of the warp/ET kernel alone, on the engine's own (interleaved, bf16) features
                prec = eng._precision(model, "fpn")
                npass = {"fp32": 0, "3xtf32": 3, "tf32": 1, "3xbf16": 3, "2xfp16": 2}[prec]
                with torch.no_grad():
                    pyr = fpn_engine.run_fpn(eng.fpn_weights, torch.cat(imgs, 0).contiguous(), npass, gen=3 if prec in ("3xbf16", "2xfp16") else 2)
                per_stage = []
                for k in range(4):
                    f = capi.cast_bf16(pyr[f"stage{k + 1}"])
                    feats = [f[v:v + 1] for v in range(nv)]
                    hypo, pose = out[f"stage{k + 1}"]["hypo_depth"], capi.pose(proj[f"stage{k + 1}"])
                    ts = []
                    for i in range(13):
                        flush.fill_(1.0)
                        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        s.record()
                        capi.et_fuse_bf16(feats[0], feats[1:], pose, hypo, G_K[k], 2.0, interleaved=bool(eng.interleave[k]))
                        e.record()
                        torch.cuda.synchronize()
                        if i >= 3:
                            ts.append(s.elapsed_time(e))
                    h, w = H >> (3 - k), W >> (3 - k)
                    nbytes = nv * C_K[k] * h * w * 2 + D_K[k] * h * w * 4 + G_K[k] * D_K[k] * h * w * 2  # bf16 features + fp32 hypotheses + bf16 cost
                    t = statistics.mean(ts)
                    per_stage.append({"stage": k + 1, "us": t * 1e3, "bytes": nbytes, "gbs": nbytes / (t * 1e-3) / 1e9, "kernel": capi.et_last_kernel()})
                tot_b, tot_t = sum(p["bytes"] for p in per_stage), sum(p["us"] for p in per_stage) * 1e-6
                res["roofline"] = {"kernel": per_stage[3]["kernel"] + f" (stage 4 launch, {H * W} pixels, {nv - 1} source views)", "bound": "hbm",
                                   "achieved": per_stage[3]["gbs"], "peak": peak, "unit": "GB/s", "frac": per_stage[3]["gbs"] / peak,
                                   "peak_source": peak_src, "algorithmic_bytes": per_stage[3]["bytes"], "traffic": None,
                                   "all_stages": {"achieved": tot_b / tot_t / 1e9, "frac": tot_b / tot_t / 1e9 / peak, "bytes": tot_b, "us": tot_t * 1e6},
                                   "per_stage": per_stage}
                del pyr, out
            del imgs, proj, dv
            eng._graphs.clear()
            torch.cuda.empty_cache()
        if ref_fp32 is not None and inputs_cfg2 is not None:  # comparison with the bf16 oracle at cfg2's size (the oracle as CHECKER)
            from oracle import mvster_oracle as oracle
            from oracle.compare import cascade_parity
            imgs_h, proj_h, dv_h, sd, cfg = inputs_cfg2
            model.storage = "bf16"
            with torch.no_grad():
                ours = model([t.to(dev) for t in imgs_h], {k: v.to(dev) for k, v in proj_h.items()}, dv_h.to(dev))
            want = oracle.cascade_forward(sd, cfg, imgs_h, proj_h, dv_h, storage_dtype=torch.bfloat16, storage_fpn_internal=False)
            rep = cascade_parity(ours, want, tie_gap=0.1, max_bad=5e-2, max_attn1=1.0)
            dist_ours = [(ours[f"stage{s}"]["attn_weight"].cpu() - want[f"stage{s}"]["attn_weight"]).abs().mean().item() for s in range(1, 5)]
            dist_fp32 = [(ref_fp32[f"stage{s}"]["attn_weight"] - want[f"stage{s}"]["attn_weight"]).abs().mean().item() for s in range(1, 5)]
            rep["mean_attn_distance_to_bf16_oracle"] = {"ours": dist_ours, "fp32_network": dist_fp32}
            rep["ok"] = bool(rep["ok"] and all(a <= 1.25 * b for a, b in zip(dist_ours, dist_fp32)))
            rep["criterion"] += "; per stage mean |attn - attn_bf16_oracle| at most 1.25x the fp32 network's distance to the same oracle (bf16 rounding is chaotic end to end: two correct implementations differ by about as much as either differs from fp32; the per-layer tests are the sharp ones)"
            res["parity"] = rep
            eng._graphs.clear()
    finally:
        model.storage = prev
    return res


def et_traffic_record(kernel_key: str):
    """DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum per launch) of the stage-4 warp/ET launch, read from the
    committed summary of an `ncu --set full` capture (profiles/et_traffic.json, written by tools/ncu_summary.py --traffic);
    None when no capture of the current kernel is on record - never a literal in this file."""
    p = REPO / "profiles" / "et_traffic.json"
    if not p.exists():
        return None, None
    import hashlib
    raw = p.read_bytes()
    try:
        rec = json.loads(raw).get(kernel_key)
    except Exception:
        return None, None
    if not rec:
        return None, None
    return float(rec["dram_bytes_read"]) + float(rec["dram_bytes_write"]), \
        {"file": "profiles/et_traffic.json", "sha256_16": hashlib.sha256(raw).hexdigest()[:16], "capture": rec.get("capture")}


def _timed_ranks(fn, steps, flush, dist, dev, warmup=3):
    """Median step time in ms over ``steps`` calls (CUDA events per step, L2 flushed between steps), max over ranks.  The
    median, because these informational legs are short and a single host hiccup (allocator, NCCL watchdog) would skew a mean."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()  # drain first: the barrier's NCCL kernel must not spin next to queued steps (see main().barrier)
    dist.barrier()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for s, e in evs:
        flush.fill_(1.0)
        s.record()
        fn()
        e.record()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    ts = [s.elapsed_time(e) for s, e in evs]
    _timed_ranks.last = [round(t, 4) for t in ts]  # this rank's step times, for the record
    total = torch.tensor([statistics.median(ts)], dtype=torch.float64, device=dev)
    dist.all_reduce(total, op=dist.ReduceOp.MAX)
    return float(total.item())


def multi_gpu_legs(args, rank, world, dev, flush):
    """The two multi-GPU configurations BASELINE.json names, measured inside the same `bench.py --gpus N` invocation:

    view_sharded - cfg4: 9 source views at 512x640, sharded over P = min(N, 4) ranks per frame (N/P frames as replicas): every
      rank extracts features for the reference view + its own source views, the warp/ET kernel writes un-normalised
      [acc | wsum] partials, ONE NCCL all-reduce per cascade stage merges them (mvs4net_utils.py:1025 is a serial loop over
      views, :1054-1060 the sum), the regulariser and head run replicated.  Reported next to the same frame unsharded on one
      GPU (strong scaling over views) and the all-reduce alone per stage.
    cfg5 - 11 views at 1024x1920, one frame per GPU (batch sharding, no collective)."""
    import torch.distributed as dist
    from mvster_b200 import sharding, synth
    steps = max(5, min(args.steps, 10))
    P = min(world, 4)
    frames = world // P
    NV, H, W = 10, 512, 640
    model = build_model(dev)
    model.use_cuda_graph = os.environ.get("MVSTER_CUDA_GRAPH", "1") == "1"
    imgs, proj, dv = synth.make_inputs(1, NV, H, W, seed=rank // P)
    imgs = [t.to(dev) for t in imgs]
    proj = {k: v.to(dev) for k, v in proj.items()}
    dv = dv.to(dev)

    def fwd():
        with torch.no_grad():
            return model(imgs, proj, dv)

    ms_full = _timed_ranks(fwd, steps, flush, dist, dev)     # P = 1: every rank runs the whole 10-view frame alone
    full_steps = list(_timed_ranks.last)
    full = fwd()
    full_depth, full_attn1 = full["depth"].clone(), full["stage1"]["attn_weight"].clone()
    shard = sharding.make_view_shard(NV - 1, P)
    model.set_view_shard(shard)
    ms_shard = _timed_ranks(fwd, steps, flush, dist, dev)
    shard_steps = list(_timed_ranks.last)
    part = fwd()
    torch.cuda.synchronize()
    same = torch.tensor([(part["depth"] == full_depth).float().mean().item(),
                         -(part["stage1"]["attn_weight"] - full_attn1).abs().max().item()], dtype=torch.float64, device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    model.set_view_shard(None)
    # the collective alone, per stage, on the same process group and payload
    ar_us, ar_bytes = [], []
    for k in range(4):
        n = sharding.allreduce_bytes(1, D_K[k], H >> (3 - k), W >> (3 - k), G_K[k])
        buf = torch.zeros(n // 4, dtype=torch.float32, device=dev)
        ms = _timed_ranks(lambda: dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=shard.group), 20, flush, dist, dev)
        ar_us.append(ms * 1e3)
        ar_bytes.append(n)
    view_sharded = {
        "workload": f"cfg4: 9 source views, 512x640, fp32; {P} view ranks per frame x {frames} frame replica(s)",
        "view_parallel": P, "frames": frames, "value": frames / (ms_shard * 1e-3), "unit": UNIT, "ms_per_step": ms_shard,
        "unsharded_ms_per_step": ms_full, "strong_scaling_vs_P1": ms_full / ms_shard,
        "rank0_step_ms": {"unsharded": full_steps, "sharded": shard_steps},
        "views_per_rank": [c for _, c in sharding.partition_views(NV - 1, P)],
        "allreduce_us_per_stage": ar_us, "allreduce_bytes": ar_bytes,
        "depth_equal_unsharded": float(same[0].item()), "stage1_attn_abs_vs_unsharded": -float(same[1].item()),
        "cuda_graph": bool(model.use_cuda_graph) and os.environ.get("MVSTER_SHARD_GRAPH", "1") == "1",
        "collective": "torch.distributed all_reduce(sum) on [acc|wsum], NCCL, one per stage"}
    del imgs, proj, dv, full, part
    eng = model._engines.get(dev.index)
    if eng is not None:
        eng._graphs.clear()
    torch.cuda.empty_cache()
    # cfg5: 11 views, 1024x1920, one frame per GPU
    imgs, proj, dv = synth.make_inputs(1, 11, 1024, 1920, seed=rank)
    imgs = [t.to(dev) for t in imgs]
    proj = {k: v.to(dev) for k, v in proj.items()}
    dv = dv.to(dev)
    ms5 = _timed_ranks(fwd, steps, flush, dist, dev)
    cfg5 = {"workload": f"cfg5: 11 views, 1024x1920 (tanks.py:58-59 crop of 1080x1920), fp32, batch {world} = one frame per GPU, no collective",
            "value": world / (ms5 * 1e-3), "unit": UNIT, "ms_per_step": ms5}
    return {"view_sharded": view_sharded, "cfg5": cfg5}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_eager"],
                    help="reference = the reference algorithm on the host CPU (oracle port); torch_eager = the same "
                         "PyTorch-op formulation the reference executes (mvster_b200/torch_path.py, eval, no_grad) as eager "
                         "CUDA kernels through cuDNN/ATen - the 'stock library kernels' bar of SURVEY.md 2b (informational)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-baseline-steps", type=int, default=3)
    ap.add_argument("--profile-range", action="store_true",
                    help="wrap the resident timed loop in cudaProfilerStart/Stop (use with ncu --profile-from-start off)")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: skip the host-buffer loop")
    ap.add_argument("--quick", action="store_true", help="skip the informational legs (GPU-eager baseline, other sizes, multi-GPU view sharding)")
    ap.add_argument("--view-parallel", type=int, default=1,
                    help="P ranks cooperate on one frame by sharding its source views (one all-reduce per stage); "
                         "world/P frames run as replicas.  Default 1 = pure batch sharding (weak scaling)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    from mvster_b200 import _lib, capi, synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = True           # as test_mvs4.py:20
    torch.backends.cudnn.allow_tf32 = False          # fp32 config: keep the feature net in true fp32
    torch.backends.cuda.matmul.allow_tf32 = False

    B, NV, H, W = WORKLOAD["batch_per_gpu"], WORKLOAD["views"], WORKLOAD["H"], WORKLOAD["W"]
    model = build_model(dev)
    # the benchmark replays the forward as one CUDA graph (config.engine.cuda_graph); MVSTER_CUDA_GRAPH=0 runs it eagerly
    model.use_cuda_graph = os.environ.get("MVSTER_CUDA_GRAPH", "1") == "1"
    P = max(1, args.view_parallel)
    if world % P:
        raise SystemExit(f"--view-parallel {P} does not divide the world size {world}")
    if P > 1:  # ranks r*P .. r*P+P-1 share one frame: each aggregates its slice of the source views,
        from mvster_b200 import sharding  # one NCCL all-reduce of [acc|wsum] per stage merges them
        model.set_view_shard(sharding.make_view_shard(NV - 1, P))
    frames = world // P
    imgs_h, proj_h, dv_h = synth.make_inputs(B, NV, H, W, seed=rank // P)
    imgs_p = [t.pin_memory() for t in imgs_h]
    proj_p = {k: v.pin_memory() for k, v in proj_h.items()}
    dv_p = dv_h.pin_memory()
    imgs_d = [t.to(dev) for t in imgs_h]
    proj_d = {k: v.to(dev) for k, v in proj_h.items()}
    dv_d = dv_h.to(dev)
    depth_host = torch.empty((B, H, W), dtype=torch.float32).pin_memory()
    conf_host = torch.empty((B, H, W), dtype=torch.float32).pin_memory()
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def step_resident():
        with torch.no_grad():
            return model(imgs_d, proj_d, dv_d)

    if args.impl == "torch_eager":  # eager ATen/cuDNN execution of the reference's op sequence on this GPU
        torch.backends.cudnn.allow_tf32 = True   # PyTorch defaults, as the reference would run
        for _ in range(max(args.warmup, 5)):
            with torch.no_grad():
                model._forward_autograd(imgs_d, proj_d, dv_d)
        torch.cuda.synchronize()
        evs = []
        for _ in range(args.steps):
            flush.fill_(1.0)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            with torch.no_grad():
                model._forward_autograd(imgs_d, proj_d, dv_d)
            e.record()
            evs.append((s, e))
        torch.cuda.synchronize()
        ms = sum(s.elapsed_time(e) for s, e in evs) / args.steps
        if rank == 0:
            print(json.dumps({"impl": "torch_eager_cuda", "metric": METRIC, "value": B / (ms * 1e-3), "unit": UNIT, "n_gpus": 1,
                              "steps": args.steps, "ms_per_step": ms, "dtype": "f32 (cuDNN TF32 allowed, PyTorch default)",
                              "note": "PyTorch-op formulation of the reference forward, eager on this GPU; informational"}))
        return

    def step_e2e():
        with torch.no_grad():
            imgs = [t.to(dev, non_blocking=True) for t in imgs_p]
            proj = {k: v.to(dev, non_blocking=True) for k, v in proj_p.items()}
            out = model(imgs, proj, dv_p.to(dev, non_blocking=True))
            depth_host.copy_(out["depth"], non_blocking=True)
            conf_host.copy_(out["photometric_confidence"], non_blocking=True)
        return out

    def drain():
        """Wait for the device.  MVSTER_BENCH_SYNC=sleep polls an event with short sleeps instead of spinning in
        cudaStreamSynchronize (A/B of host-side interference at 8 ranks per node)."""
        if os.environ.get("MVSTER_BENCH_SYNC", "spin") == "sleep":
            ev = torch.cuda.Event()
            ev.record()
            while not ev.query():
                time.sleep(0.0002)
        else:
            torch.cuda.synchronize()

    def barrier():
        """Rendezvous of the ranks around a timed region.  The device is drained BEFORE the NCCL barrier: ProcessGroupNCCL runs the
        barrier's all-reduce on its own stream, where it would spin on a few SMs next to the still queued steps - and a persistent
        convolution kernel that finds 140 of 148 SMs free needs two waves (measured at 8 ranks: steps at 2.5-8.6 ms instead of 1.58
        once the host had enqueued everything and reached the barrier)."""
        drain()
        if world > 1:
            dist.barrier()
            drain()

    def timed_e2e_pipelined(steps, inflight=1):
        """K end-to-end steps as a serving loop runs them: the pinned-host -> device copy of step i+1's inputs is enqueued on a
        copy stream (ring of device input buffers) while step i's forward runs; depth + confidence go back to pinned host
        memory after every forward.  With ``inflight`` = 2 consecutive frames alternate between two compute streams and two graph
        slots (``model.graph_slot``), so that one frame's feature pyramid overlaps the other's cascade stages.  Every step's H2D,
        L2 flush, forward and D2H lie inside the one timed region (CUDA events on the launching stream around all K steps).
        Returns the elapsed ms (max over ranks)."""
        main, copy_stream = torch.cuda.current_stream(dev), torch.cuda.Stream(device=dev)
        cstreams = [main] if inflight == 1 else [torch.cuda.Stream(device=dev) for _ in range(inflight)]
        nbuf = inflight + 1
        bufs = [{"imgs": [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in imgs_p],
                 "proj": {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in proj_p.items()},
                 "dv": torch.empty(dv_p.shape, dtype=dv_p.dtype, device=dev),
                 "ready": torch.cuda.Event(), "free": torch.cuda.Event()} for _ in range(nbuf)]
        hosts = [(depth_host, conf_host)] + [(torch.empty_like(depth_host).pin_memory(), torch.empty_like(conf_host).pin_memory())
                                              for _ in range(inflight - 1)]

        def h2d(i):
            b = bufs[i % nbuf]
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(b["free"])  # the forward that read this buffer (step i - nbuf) has consumed it
                for dst, src in zip(b["imgs"], imgs_p):
                    dst.copy_(src, non_blocking=True)
                for k, dst in b["proj"].items():
                    dst.copy_(proj_p[k], non_blocking=True)
                b["dv"].copy_(dv_p, non_blocking=True)
                b["ready"].record(copy_stream)

        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        s.record(main)
        for cs in cstreams:
            cs.wait_stream(main)
        for i in range(min(inflight, steps)):
            h2d(i)
        slot0 = model.graph_slot
        with torch.no_grad():
            for i in range(steps):
                if i + inflight < steps:
                    h2d(i + inflight)
                b, cs = bufs[i % nbuf], cstreams[i % inflight]
                dh, ch = hosts[i % inflight]
                with torch.cuda.stream(cs):
                    cs.wait_event(b["ready"])
                    flush.fill_(1.0)  # the L2 flush stays inside the timed loop here (conservative)
                    model.graph_slot = i % inflight
                    out = model(b["imgs"], b["proj"], b["dv"])
                    b["free"].record(cs)
                    dh.copy_(out["depth"], non_blocking=True)
                    ch.copy_(out["photometric_confidence"], non_blocking=True)
        model.graph_slot = slot0
        for cs in cstreams:
            main.wait_stream(cs)
        e.record(main)
        barrier()
        if inflight > 1 and steps > 0:  # the loop's last frame sits in the last slot's host buffers
            depth_host.copy_(hosts[(steps - 1) % inflight][0])
        total = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(total, op=dist.ReduceOp.MAX)
        return float(total.item())

    def timed(fn, steps):
        """K steps, each bracketed by CUDA events on the launching stream; L2 flushed between steps
        (outside the event pairs).  Returns the summed step time in ms (max over ranks)."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for s, e in evs:
            flush.fill_(1.0)
            s.record()
            fn()
            e.record()
        barrier()
        per_step.clear()
        per_step.extend(s.elapsed_time(e) for s, e in evs)
        total = torch.tensor([sum(per_step)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(total, op=dist.ReduceOp.MAX)
        return float(total.item())

    per_step = []  # this rank's step times (ms) of the last timed() call
    # kernels of libmvster_b200 per forward, counted on one eager pass (a CUDA-graph replay issues the same kernels
    # without going through the library's launch counter)
    graphed, model.use_cuda_graph = model.use_cuda_graph, False
    lc0 = _lib.launch_count()
    step_resident()
    launches_per_step = _lib.launch_count() - lc0
    model.use_cuda_graph = graphed
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # before the warm-up: spawning nvidia-smi from a process this size stalls the host for tens of ms
    for _ in range(args.warmup):
        step_resident()
    for _ in range(2):
        step_e2e()
    l0 = _lib.launch_count()
    if args.profile_range:
        torch.cuda.profiler.start()
    import gc
    gc.collect()
    ms_total = timed(step_resident, args.steps)
    step_stats = {"min_ms": min(per_step), "median_ms": statistics.median(per_step),
                  "p95_ms": sorted(per_step)[min(len(per_step) - 1, int(0.95 * len(per_step)))], "max_ms": max(per_step),
                  "steps_ms": [round(t, 3) for t in per_step]}
    if args.profile_range:
        torch.cuda.profiler.stop()
    launches = launches_per_step * args.steps if (model.use_cuda_graph and P == 1) else _lib.launch_count() - l0
    ms_e2e_serial = ms_total if args.skip_e2e else timed(step_e2e, args.steps)
    ms_e2e, e2e_pipelined, inflight, ms_e2e_1, ms_e2e_n = ms_e2e_serial, False, 1, None, None
    if not args.skip_e2e and P == 1 and os.environ.get("MVSTER_BENCH_E2E", "pipelined") == "pipelined":
        e2e_pipelined = True
        timed_e2e_pipelined(3)  # warm-up of the copy stream and the double buffers
        ms_e2e = timed_e2e_pipelined(args.steps)
        # MVSTER_BENCH_INFLIGHT=2: additionally run the loop with two frames in flight (two graph slots, two compute streams) and take
        # the better one.  Off by default: measured 650-672 maps/s against 567 in some runs and 159-353 in others - the two graphs'
        # persistent kernels each size their grid for the whole GPU and interleave differently from run to run
        inflight = int(os.environ.get("MVSTER_BENCH_INFLIGHT", "1"))
        if inflight > 1:
            timed_e2e_pipelined(2 * inflight, inflight)  # captures the second slot's graph, warms the streams
            ms_e2e_1, ms_e2e_n = ms_e2e, timed_e2e_pipelined(args.steps, inflight)
            ms_e2e = min(ms_e2e_1, ms_e2e_n)
        torch.cuda.synchronize()
        with torch.no_grad():  # the pipelined loop must deliver the same frame as the resident call
            ref_out = step_resident()
            torch.cuda.synchronize()
            if not torch.equal(depth_host, ref_out["depth"].cpu()):
                raise SystemExit("bench.py: pipelined end-to-end loop returned a different depth map than the resident call")
    clocks = sampler.stop() if rank == 0 else None
    ms_step, ms_step_e2e = ms_total / args.steps, ms_e2e / args.steps
    value = frames * B / (ms_step * 1e-3)
    e2e_value = frames * B / (ms_step_e2e * 1e-3)

    # ---- warp (ET) kernel alone: live CUDA-event timing per stage, L2 flushed between launches
    peak, peak_src = measured_peaks()
    roof, breakdown = None, None
    out = step_resident()  # every rank: under --view-parallel the forward holds a collective
    if rank == 0:
        with torch.no_grad():
            # the features exactly as the engine hands them to the kernel: its own native pyramid (stages 2-4 group-interleaved)
            from mvster_b200 import fpn_engine
            eng = model._engines[dev.index]
            prec = eng._precision(model, "fpn")
            npass = {"fp32": 0, "3xtf32": 3, "tf32": 1, "3xbf16": 3, "2xfp16": 2}[prec]
            pyr = fpn_engine.run_fpn(eng.fpn_weights, torch.cat(imgs_d, 0).contiguous(), npass, gen=3 if prec in ("3xbf16", "2xfp16") else 2)
            per_stage = []
            for k in range(4):
                f = pyr[f"stage{k + 1}"]
                il = bool(eng.interleave[k])
                feats = [f[v * B:(v + 1) * B] for v in range(NV)]
                hypo = out[f"stage{k + 1}"]["hypo_depth"]
                pose = capi.pose(proj_d[f"stage{k + 1}"])
                cost = torch.empty((B, D_K[k], H >> (3 - k), W >> (3 - k), G_K[k]), device=dev)
                for _ in range(3):
                    capi.et_fuse(feats[0], feats[1:], pose, hypo, G_K[k], 2.0, cost=cost, interleaved=il)
                ts = []
                for _ in range(20):
                    flush.fill_(1.0)
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record()
                    capi.et_fuse(feats[0], feats[1:], pose, hypo, G_K[k], 2.0, cost=cost, interleaved=il)
                    e.record()
                    torch.cuda.synchronize()
                    ts.append(s.elapsed_time(e))
                t = statistics.mean(ts)
                nbytes = et_algorithmic_bytes(k, B, NV - 1, H, W)
                per_stage.append({"stage": k + 1, "us": t * 1e3, "bytes": nbytes, "gbs": nbytes / (t * 1e-3) / 1e9})
        tot_b = sum(p["bytes"] for p in per_stage)
        tot_t = sum(p["us"] for p in per_stage) * 1e-6
        dom = per_stage[3]
        kname = capi.et_last_kernel()  # the stage-4 launch was the last one timed
        traffic, traffic_src = et_traffic_record(kname) if (B, NV, H, W) == (1, 5, 512, 640) else (None, None)
        roof = {"kernel": kname + " (stage 4 launch: C=8, G=4, D=4, 4 source views, 327680 pixels)", "bound": "hbm",
                "achieved": dom["gbs"], "peak": peak, "unit": "GB/s", "frac": dom["gbs"] / peak, "peak_source": peak_src,
                "algorithmic_bytes": dom["bytes"], "traffic": traffic, "traffic_source": traffic_src,
                "all_stages": {"achieved": tot_b / tot_t / 1e9, "frac": tot_b / tot_t / 1e9 / peak, "bytes": tot_b, "us": tot_t * 1e6},
                "per_stage": per_stage}

    # ---- the convolution kernel that holds most of the step time (profiles/r01_launches_tc3.md): tensor-pipe view.
    # One launch of reg2d conv2 at stage 4 (16 -> 16 channels, 3x3x3, D x H/2 x W/2 voxels), timed alone like the ET launches.
    roof_tc = None
    if rank == 0 and model.reg_precision in ("3xbf16", "2xfp16"):
        from mvster_b200 import packing
        tf_peak = tensor_peak()
        split = 2 if model.reg_precision == "2xfp16" else 3
        vox = (B, D_K[3], H // 2, W // 2)
        xin = torch.randn(*vox, 16, device=dev)
        wp = packing.pack_tc3_weights(torch.randn(27, 16, 16) / 20.0, 3, 3, 1, split).to(dev)
        bias = torch.zeros(16, device=dev)
        # the form the engine runs by default: two fp16 terms with PACKED operands (activations stored as the (a1, a2) planes the
        # producing layer's epilogue wrote; MVSTER_REG_PACKED=0 or three bf16 terms: fp32 activations + converter warps)
        packed_tc = split == 2 and os.environ.get("MVSTER_REG_PACKED", "1") == "1"
        if packed_tc:
            xp = capi.to_ph16(xin)
            layer = lambda: capi.conv_tc3_pb16(xp, wp, bias, 16, 3, 3, 1, True)
        else:
            yout = torch.empty(*vox, 16, device=dev)
            layer = lambda: capi.conv_tc3(xin, wp, bias, 16, 3, 3, 1, True, out=yout, split=split)
        for _ in range(3):
            layer()
        ts = []
        for _ in range(20):
            flush.fill_(1.0)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            layer()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        t = statistics.mean(ts)
        flops = 2.0 * vox[0] * vox[1] * vox[2] * vox[3] * 27 * 16 * 16
        roof_tc = {"kernel": "conv_tc3_kernel<16,%d%s> (reg2d conv2 at stage 4: 16 -> 16 channels, 3x3x3, %d voxels)" % (split, ", packed operands" if packed_tc else "", vox[0] * vox[1] * vox[2] * vox[3]),
                   "bound": "tensor", "achieved": flops / (t * 1e-3) / 1e12, "peak": tf_peak[0], "unit": "TFLOP/s",
                   "frac": flops / (t * 1e-3) / 1e12 / tf_peak[0], "peak_source": tf_peak[1], "us": t * 1e3,
                   "algorithmic_flops": flops,
                   "note": ("fp32-faithful conv flops; the kernel issues 3 fp16 products per fp32 product (2 MMAs of N = 32/16 per tap and 16 channels)"
                            if split == 2 else
                            "fp32-faithful conv flops; the kernel issues 6 bf16 products per fp32 product (3 MMAs of N = 48/32/16 per tap and 16 channels)") +
                           "; every M128 x K16 MMA is bound by its 4 KB A-operand fetch from shared memory, not by the multipliers, and the"
                           " layer by the activation ring's latency (profiles/r01_conv_tc3_h16_ncu.md)", "traffic": None}

    cpu_base, parity, bf16_inputs, ref_out = None, None, None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import mvster_oracle as oracle
        from oracle.compare import cascade_parity
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        cfg = dict(oracle.DEFAULT_CFG)
        small = synth.make_inputs(1, 3, 128, 192, seed=1)
        cores = pick_cpu_threads(lambda: oracle.cascade_forward(sd, cfg, *small))
        ref_out = oracle.cascade_forward(sd, cfg, imgs_h, proj_h, dv_h)
        t0 = time.perf_counter()
        for _ in range(args.cpu_baseline_steps):
            oracle.cascade_forward(sd, cfg, imgs_h, proj_h, dv_h)
        dt = (time.perf_counter() - t0) / args.cpu_baseline_steps
        cpu_base = {"value": B / dt, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"{args.cpu_baseline_steps} full forwards of the same frame after 1 warm-up (oracle port, torch CPU fp32)"}
        # the oracle as CHECKER: the frame the timed loop produced (same inputs, same weights) against the oracle's
        parity = cascade_parity(step_resident(), ref_out)
        bf16_inputs = (imgs_h, proj_h, dv_h, sd, cfg)

    eager, extra = None, None
    if rank == 0 and world == 1 and not args.quick:
        eager = gpu_eager_baseline(model, imgs_d, proj_d, dv_d, flush)
        extra = extra_legs(model, dev, flush)
        graphed, model.use_cuda_graph = model.use_cuda_graph, False  # cfg2 through eager launches (no CUDA graph)
        ms, best = _time_forward(step_resident, 10, flush)
        model.use_cuda_graph = graphed
        extra["no_cuda_graph"] = {"ms_per_step": ms, "min_ms": best, "value": B / (ms * 1e-3)}
        try:  # throughput with two frames per forward (batch 2 in ONE graph): the latency-bound early stages do twice the work per launch
            i2, p2, d2 = synth.make_inputs(2, NV, H, W, seed=0)
            i2, p2, d2 = [t.to(dev) for t in i2], {k: v.to(dev) for k, v in p2.items()}, d2.to(dev)

            def fwd2():
                with torch.no_grad():
                    return model(i2, p2, d2)
            ms2, best2 = _time_forward(fwd2, 10, flush)
            extra["batch2"] = {"ms_per_step": ms2, "min_ms": best2, "value": 2.0 / (ms2 * 1e-3), "unit": UNIT,
                               "note": "informational: two frames per forward; the headline stays one frame per forward (the reference's test_mvs4.py batch size)"}
            del i2, p2, d2
            model._engines[dev.index]._graphs.clear()
            torch.cuda.empty_cache()
        except Exception as exc:  # noqa: BLE001
            extra["batch2"] = {"error": f"{type(exc).__name__}: {exc}"[:400]}
        try:  # informational: a failure here must not take the headline down with it
            extra["bf16_storage"] = bf16_storage_leg(model, dev, flush, ref_out, bf16_inputs)
        except Exception as exc:  # noqa: BLE001
            extra["bf16_storage"] = {"error": f"{type(exc).__name__}: {exc}"[:400]}

    multi = None
    if world > 1 and not args.quick and P == 1:
        multi = multi_gpu_legs(args, rank, world, dev, flush)

    if rank == 0:
        h2d = sum(t.numel() * 4 for t in imgs_p) + sum(v.numel() * 4 for v in proj_p.values()) + dv_p.numel() * 4
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak" if P == 1 else f"weak over {frames} frame(s), strong x{P} over source views",
            "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": dict(config_dict(world, P), engine={
                "fpn_backend": model.fpn_backend, "fpn_precision": model.fpn_precision, "reg_precision": model.reg_precision,
                "reg_packed_operands": os.environ.get("MVSTER_REG_PACKED", "1") == "1" and model.reg_precision == "2xfp16",
                "cuda_graph": bool(model.use_cuda_graph) and P == 1,
                "overlap_stages": bool(getattr(model, "overlap_stages", False)) and P == 1}),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_step_e2e,
                    "mode": "serving loop: step i+1's pinned-host -> device copy on a copy stream (double-buffered inputs) under step i's "
                            "forward; L2 flush, forward and D2H of every step inside the one timed region; the better of 1 and "
                            f"{inflight} frames in flight (model.graph_slot: one captured graph per slot, alternating compute streams)" if e2e_pipelined else
                            "serial: H2D, forward, D2H on one stream, each step bracketed by CUDA events",
                    "serial_value": frames * B / (ms_e2e_serial / args.steps * 1e-3), "h2d_bytes_per_step": h2d,
                    **({"frames_in_flight": {"1": frames * B / (ms_e2e_1 / args.steps * 1e-3), str(inflight): frames * B / (ms_e2e_n / args.steps * 1e-3)}}
                       if e2e_pipelined and inflight > 1 else {}),
                    "d2h_bytes_per_step": 2 * B * H * W * 4},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "roofline_tensor": roof_tc, "cpu_baseline": cpu_base,
            "step_stats": step_stats, "parity": parity, "gpu_eager_baseline": eager, "extra": extra,
            **({} if multi is None else multi),
        }))
        if parity is not None and not parity["ok"]:
            sys.stdout.flush()
            raise SystemExit("bench.py: the benchmarked frame does not match the oracle: " + json.dumps(parity))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
