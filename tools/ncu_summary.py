#!/usr/bin/env python
"""Condense an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the few rows the roofline
discussion needs, as a markdown table.  Usage: tools/ncu_summary.py report.ncu-rep [kernel-regex]"""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1TEX throughput % (active)"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data-pipe wavefronts %"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
]


def traffic(rep, pattern, key, out):
    """--traffic REPORT KERNEL_REGEX KEY OUT.json: DRAM bytes per launch (mean over the matching launches) of an `ncu --set full`
    capture, stored under KEY (the kernel name bench.py prints, capi.et_last_kernel()) - bench.py reads roofline.traffic from it."""
    import json
    import os
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    pat = re.compile(pattern)
    sel = [r for r in data if pat.search(r[idx["Kernel Name"]])]
    if not sel:
        raise SystemExit(f"no launch matches {pattern!r}")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def mean(metric):
        u = scale[units[idx[metric]]]
        return sum(float(r[idx[metric]].replace(",", "")) for r in sel) / len(sel) * u
    rec = {"dram_bytes_read": mean("dram__bytes_read.sum"), "dram_bytes_write": mean("dram__bytes_write.sum"), "launches": len(sel),
           "duration_us_under_ncu": sum(float(r[idx["gpu__time_duration.sum"]].replace(",", "")) for r in sel) / len(sel)
                                    * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[idx["gpu__time_duration.sum"]].replace("second", "s").replace("usecond", "us"), 1.0),
           "capture": os.path.basename(rep), "kernel_regex": pattern}
    cur = json.load(open(out)) if os.path.exists(out) else {}
    cur[key] = rec
    json.dump(cur, open(out, "w"), indent=1, sort_keys=True)
    print(json.dumps({key: rec}, indent=1))


def main():
    if sys.argv[1] == "--traffic":
        return traffic(*sys.argv[2:6])
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in data if pat is None or pat.search(r[idx["Kernel Name"]])]
    names = [re.sub(r"^void ", "", r[idx["Kernel Name"]])[:60] for r in data]
    print("| metric | " + " | ".join(names) + " |")
    print("|---|" + "---|" * len(names))
    for key, label in KEYS:
        if key not in idx:
            continue
        vals = []
        for r in data:
            v = r[idx[key]]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:,.0f}" if abs(f) >= 1000 else f"{f:.3g}"
            except ValueError:
                pass
            vals.append(f"{v} {units[idx[key]]}".strip())
        print(f"| {label} (`{key}`) | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
