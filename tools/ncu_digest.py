"""Digest of an .ncu-rep file read on the CPU box: key metrics per kernel, stall reasons, opcode histogram and the
hottest SASS lines.   usage: python tools/ncu_digest.py file.ncu-rep [kernel-index] [n-hot-lines]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
nhot = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.sum", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w} = {r[i]} {units[i]}")
    for i, h in enumerate(hdr):
        if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(r[i] or 0) > 0.05:
            print("   stall", h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), r[i])
    print()
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
kern, data = None, collections.OrderedDict()
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        kern = r[1] + f"#{len(data)}"; data[kern] = []; continue
    if r and r[0] == "Address":
        shdr = r; continue
    if kern and len(r) > 6:
        data[kern].append(r)
k = list(data)[kidx]
rs = data[k]
iE, iS = shdr.index("Instructions Executed"), shdr.index("# Samples")
tot = sum(int(r[iE]) for r in rs); tots = sum(int(r[iS]) for r in rs)
print(k, "instructions", tot, "samples", tots)
op, ops = collections.Counter(), collections.Counter()
for r in rs:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[1]); o = m.group(2).split(".")[0] if m else "?"
    op[o] += int(r[iE]); ops[o] += int(r[iS])
for o, c in op.most_common(22):
    print(f"   {o:10s} {c / tot * 100:5.1f}% of instructions  {ops[o] / max(tots,1) * 100:5.1f}% of samples")
print("hottest lines by samples:")
stall_cols = [i for i, h in enumerate(shdr) if h.startswith("stall_") and "Not Issued" not in h]
for i, r in sorted(enumerate(rs), key=lambda t: -int(t[1][iS]))[:nhot]:
    st = sorted(((int(r[c] or 0), shdr[c]) for c in stall_cols), reverse=True)[:2]
    print(f"  {i:5d} samples {int(r[iS]):5d} exec {int(r[iE]):9d}  {r[1].strip():60s} {st}")
