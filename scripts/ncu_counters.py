"""Development aid: per-env-step instruction / FLOP counters of the world kernel from one ncu raw-page CSV
(-> profiles/world_kernel_counters.json, which bench.py's compute roofline reads)."""
import csv, json, sys
raw, worlds, source = sys.argv[1], int(sys.argv[2]), sys.argv[3]
rows = list(csv.reader(open(raw)))
d = {h: v for h, v in zip(rows[0], rows[2])}
f = lambda k: float(d[k])
cyc = f("sm__cycles_elapsed.max")
ffma, fadd, fmul = [f(f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum.per_cycle_elapsed") * cyc for k in ("ffma", "fadd", "fmul")]
warp_inst = f("smsp__inst_executed.sum")
lanes = f("smsp__thread_inst_executed_per_inst_executed.ratio")
out = {
    "source": source, "worlds_per_launch": worlds, "kernel_ms_under_ncu": f("gpu__time_duration.sum"),
    "warp_inst_per_env_step": warp_inst / worlds, "thread_inst_per_env_step": warp_inst * lanes / worlds, "lanes_per_inst": lanes,
    "fp32_flop_per_env_step": (2 * ffma + fadd + fmul) / worlds, "ffma_per_env_step": ffma / worlds, "fadd_per_env_step": fadd / worlds,
    "fmul_per_env_step": fmul / worlds, "fp32_share_of_thread_inst": (ffma + fadd + fmul) / (warp_inst * lanes),
    "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"), "ipc_per_sm": f("sm__inst_executed.avg.per_cycle_elapsed"),
    "warps_per_sm": f("sm__warps_active.avg.pct_of_peak_sustained_active") / 100 * 64,
    "dram_bytes_per_launch": (f("dram__bytes_read.sum") + f("dram__bytes_write.sum")) * 1e6,
    "local_loads_per_env_step": f("smsp__sass_inst_executed_op_local_ld.sum") / worlds, "local_stores_per_env_step": f("smsp__sass_inst_executed_op_local_st.sum") / worlds,
}
json.dump(out, open(sys.argv[4], "w"), indent=1)
print(json.dumps(out, indent=1))
