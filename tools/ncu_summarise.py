"""Turn an .ncu-rep (ncu --set full) into a compact per-launch CSV for profiles/ (run where ncu is installed).
usage: ncu_summarise.py report.ncu-rep out.csv [label ...]   -- optional labels name the launches in order"""
import csv, subprocess, sys

METRICS = [("gpu__time_duration.sum", "duration_us"),
           ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct"),
           ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
           ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_shared_pct"),
           ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
           ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
           ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
           ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
           ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
           ("launch__cluster_size", "cluster"), ("sm__cycles_elapsed.avg.per_second", "sm_clock")]
rep, out = sys.argv[1], sys.argv[2]
labels = sys.argv[3:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
head, units, data = rows[0], rows[1], rows[2:]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["launch", "kernel"] + [f"{n} [{units[head.index(m)]}]" if m in head else n for m, n in METRICS])
    for i, r in enumerate(data):
        d = dict(zip(head, r))
        w.writerow([labels[i] if i < len(labels) else i, d.get("Kernel Name", "")] + [d.get(m, "") for m, _ in METRICS])
print(f"{len(data)} launches -> {out}")
