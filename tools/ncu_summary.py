#!/usr/bin/env python
"""Condenses `ncu -i REPORT --page raw --csv` exports into the JSON kept under
profiles/ (r01_sweep_ncu_summary.json): per kernel launch the duration, DRAM
bytes, pipe and issue utilisation, occupancy limits and the stall reasons per
issued instruction.

    python tools/ncu_summary.py --out profiles/r01_sweep_ncu_summary.json \
        --s16 gpurun_out/prof_s16_raw.csv --f32 gpurun_out/prof_f32_raw.csv \
        [--more gpurun_out/prof_post_raw.csv] --alg-bytes 495190600 \
        --what "..." --command "..."
"""
from __future__ import annotations

import argparse
import csv
import json

PLAIN = [
    "dram__bytes_read.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__registers_per_thread",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active",
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"
STALL_SUFFIX = "_per_issue_active.ratio"


def _num(s: str):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def kernels_of(path: str):
    rows = list(csv.reader(open(path, newline="")))
    head, units = rows[0], rows[1]
    col = {name: i for i, name in enumerate(head)}
    out = []
    for r in rows[2:]:
        m = {}
        for name in PLAIN:
            if name in col and _num(r[col[name]]) is not None:
                m[name] = {"value": _num(r[col[name]]), "unit": units[col[name]]}
        for name, i in col.items():
            if name.startswith(STALL_PREFIX) and name.endswith(STALL_SUFFIX):
                v = _num(r[i])
                if v is not None and v >= 0.05:
                    key = "stall_" + name[len(STALL_PREFIX):-len(STALL_SUFFIX)] + "_per_issue"
                    m[key] = {"value": v, "unit": units[i]}
        out.append({"kernel": r[col["Kernel Name"]], "metrics": m})
    return out


def _bytes(metric):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return metric["value"] * scale[metric["unit"]]


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--s16", required=True)
    ap.add_argument("--f32")
    ap.add_argument("--more", nargs="*", default=[])
    ap.add_argument("--alg-bytes", type=int, required=True)
    ap.add_argument("--what", default="")
    ap.add_argument("--command", default="")
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    ks = kernels_of(a.s16)
    sweep = ks[0]["metrics"]
    rd, wr = _bytes(sweep["dram__bytes_read.sum"]), _bytes(sweep["dram__bytes_write.sum"])
    for path in a.more:
        ks += kernels_of(path)
    doc = {"what": a.what, "command": a.command, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "traffic_bytes_per_launch": rd + wr, "algorithmic_bytes_per_launch": a.alg_bytes,
           "kernels": ks, "float32_input": kernels_of(a.f32) if a.f32 else []}
    with open(a.out, "w") as f:
        json.dump(doc, f, indent=1)
    print(f"{a.out}: {len(ks)} kernel(s), sweep traffic {rd + wr:.0f} B per launch")


if __name__ == "__main__":
    main()
