"""Markdown summary of an ncu report: python tools/ncu_summary.py gpurun_out/x.ncu-rep [kernel-regex] > profiles/x.md
Reads the report with `ncu -i ... --page raw --csv` (works without a GPU)."""
import csv, io, re, subprocess, sys

rep = sys.argv[1]
kre = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
kn = hdr.index("Kernel Name")
if kre:
    data = [r for r in data if kre.search(r[kn])]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum"]
WANT += [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h]
print("| metric | " + " | ".join(f"launch {i + 1} ({r[kn][:24]})" for i, r in enumerate(data)) + " | unit |")
print("|---|" + "---|" * (len(data) + 1))
for w in WANT:
    if w not in hdr:
        continue
    i = hdr.index(w)
    vals = [r[i] for r in data]
    try:
        if all(float(v) < 0.05 for v in vals) and "stalled" in w:
            continue
    except ValueError:
        pass
    print(f"| `{w}` | " + " | ".join(vals) + f" | {units[i]} |")
