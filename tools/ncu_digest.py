#!/usr/bin/env python3
"""Digest of an ncu report: headline metrics, stall mix, hottest SASS lines, opcode mix."""
import csv, re, subprocess, sys, collections, io
rep = sys.argv[1]
frames = float(sys.argv[2]) if len(sys.argv) > 2 else None   # warp-frames, for per-frame figures
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = dict(zip(hdr, vals))
def g(k):
    return m.get(k, "?")
print("kernel", g("Kernel Name")[:60], "dur us", g("gpu__time_duration.sum"), "regs", g("launch__registers_per_thread"),
      "grid", g("launch__grid_size"), "block", g("launch__block_size"))
for k in ("sm__cycles_active.avg", "sm__cycles_elapsed.max", "sm__warps_active.avg.per_cycle_active",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
          "smsp__warps_eligible.avg.per_cycle_active"):
    print(f"  {k:75s} {g(k)}")
st = {k.split("issue_stalled_")[1].split("_per_")[0]: float(v) for k, v in m.items()
      if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v not in ("", "?")}
print("  stalls per issue:", ", ".join(f"{k}={v:.2f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v > 0.01))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]]) for r in data)
ops = collections.Counter(); samp = collections.Counter(); ninst = 0
for r in data:
    mm = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[1].strip())
    op = mm.group(2) if mm else r[1][:10]
    n = int(r[ix["Instructions Executed"]]); ops[op] += n; ninst += n; samp[op] += int(r[ix["# Samples"]])
print(f"  samples {tot}, warp instructions {ninst}" + (f", {ninst / frames:.2f} per warp-frame" if frames else ""))
print("  opcode mix:", ", ".join(f"{op} {n / (frames or ninst):.2f} ({100 * samp[op] / tot:.0f}%s)" for op, n in ops.most_common(18)))
print("  hottest lines:")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:14]:
    print(f"    {int(r[ix['# Samples']]):6d} ({100 * int(r[ix['# Samples']]) / tot:4.1f}%) x{r[ix['Instructions Executed']]:>8s}  {r[1].strip()[:80]}")
