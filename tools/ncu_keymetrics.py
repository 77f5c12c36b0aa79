"""Key metrics of one .ncu-rep (first captured launch of the named kernel) as a markdown table row set.
usage: python tools/ncu_keymetrics.py report.ncu-rep [label]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
label = sys.argv[2] if len(sys.argv) > 2 else rep
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, r = rows[0], rows[2]
d = dict(zip(hdr, r))
keys = [("gpu__time_duration.sum", "duration (us, isolated launch)"),
        ("dram__bytes_read.sum", "DRAM read (MB)"), ("dram__bytes_write.sum", "DRAM written (MB)"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy (%)"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe (%)"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy (%)"),
        ("launch__registers_per_thread", "registers / thread"),
        ("launch__shared_mem_per_block_dynamic", "dynamic shared memory / CTA (KB)"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
        ("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "local-memory load sectors"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (cycles / issue)"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle")]
print(f"### {label}\n\n| metric | value |\n|---|---|")
for k, name in keys:
    v = d.get(k)
    if v in (None, ""):
        continue
    try:
        f = float(v)
        v = f"{f:,.0f}" if abs(f) >= 1000 else f"{f:.3g}"
    except ValueError:
        pass
    print(f"| {name} (`{k}`) | {v} |")
print()
