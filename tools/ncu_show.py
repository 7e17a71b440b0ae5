"""Print the key metrics of raw-CSV ncu exports side by side.  usage: ncu_show.py a.raw.csv [b.raw.csv ...]"""
import csv, sys
KEYS = ["gpu__time_duration.sum","launch__grid_size","launch__block_size","launch__registers_per_thread",
 "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct_of_peak_sustained_active",
 "sm__warps_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
 "smsp__inst_executed.sum","smsp__thread_inst_executed_per_inst_executed.ratio","sm__cycles_elapsed.avg.per_second",
 "dram__bytes_read.sum","dram__bytes_write.sum"]
STALL = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
for s in ["wait","math_pipe_throttle","not_selected","long_scoreboard","short_scoreboard","barrier","sleeping","no_instruction","dispatch_stall","branch_resolving","membar","mio_throttle","lg_throttle","tex_throttle","drain","imc_miss","misc","selected"]:
    KEYS.append(STALL % s)
KEYS += ["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum","smsp__sass_thread_inst_executed_op_dmul_pred_on.sum","smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
tabs = []
for p in sys.argv[1:]:
    rows = [r for r in csv.reader(open(p)) if r]
    hdr, units, vals = rows[0], rows[1], rows[2]
    tabs.append({h: (v, u) for h, u, v in zip(hdr, units, vals)})
print("kernels:", [t.get("Kernel Name", ("?",))[0][:60] for t in tabs])
for k in KEYS:
    if any(k in t for t in tabs):
        print(f"{k[:78]:78s}", *[f"{t.get(k, ('-',''))[0]:>16s}" for t in tabs], tabs[0].get(k, ("", ""))[1])
