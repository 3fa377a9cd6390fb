"""Development aid: markdown summary of one ncu kernel capture (raw page CSV + optional per-function table)."""
import csv, subprocess, sys
raw_csv, title = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw_csv)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
keys = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__icc_request_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
print(f"# {title}\n\n| metric | value | unit |\n|---|---|---|")
for k in keys:
    if k in d:
        print(f"| {k} | {d[k][1]} | {d[k][0]} |")
st = {h: float(v[1]) for h, v in d.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h}
print("\nWarp stall reasons (cycles per issued instruction): " + ", ".join(f"{h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} {v:.2f}" for h, v in sorted(st.items(), key=lambda kv: -kv[1]) if v >= 0.01))
if len(sys.argv) > 4:
    out = subprocess.run([sys.executable, __file__.replace("ncu_summary", "ncu_by_function"), sys.argv[3], sys.argv[4]], capture_output=True, text=True).stdout
    print("\nPer device function (SASS source page aggregated by the noinline phase functions):\n\n```\n" + out + "```")
