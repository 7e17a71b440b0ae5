#!/usr/bin/env python
"""Turn the scratch outputs of tools/gpu_round.sh (gpurun_out/) into the tracked summaries under
profiles/.  usage: summarise_profiles.py <tag>   (e.g. r01a)

  profiles/<tag>_launches.md   per-kernel launch count / total / share from the ncu launch list
  profiles/<tag>_s1_ncu.md     key metrics of the `ncu --set full` capture of the dominant kernel
  profiles/<tag>_bench.json    the bench line of the same box session
  profiles/traffic.json        DRAM bytes per launch of the dominant kernel, scaled to the bench
                               launch (read by bench.py for roofline.traffic)
"""
import csv
import json
import os
import re
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.environ.get("SUMMARY_OUT") or os.path.join(ROOT, "profiles")

KEYS = [
    "gpu__time_duration.sum",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    # executed FP64 instruction mix, thread instructions per elapsed cycle summed over the GPU
    # (peak 64 lanes x 148 SMs = 9472): executed flop/cycle = 2*dfma + dmul + dadd
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
    "sm__inst_executed.sum.per_cycle_elapsed",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__cycles_elapsed.max",
]


def launches(tag):
    path = os.path.join(OUT, "launches.csv")
    if not os.path.exists(path):
        return None
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], float(r["Metric Value"]), r["Grid Size"], r["Block Size"]))
    agg = defaultdict(lambda: [0, 0.0, "", ""])
    for name, ns, g, b in rows:
        short = re.sub(r"\(.*", "", name)
        a = agg[short]
        a[0] += 1
        a[1] += ns
        a[2], a[3] = g, b
    total = sum(a[1] for a in agg.values()) or 1.0
    out = [
        f"# {tag}: ncu launch list of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline`",
        "",
        "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare shares).",
        "",
        "| kernel | launches | total ms | share | grid | block |",
        "|---|---|---|---|---|---|",
    ]
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{name}` | {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / total:.2f}% | {a[2]} | {a[3]} |")
    out.append("")
    out.append(f"{len(rows)} launches, {total / 1e6:.1f} ms of device time in total.")
    open(os.path.join(PROF, f"{tag}_launches.md"), "w").write("\n".join(out) + "\n")
    return agg


def ncu_full(tag, rep, title, launch_units=None):
    path = os.path.join(OUT, rep)
    csv_path = os.path.join(OUT, os.path.splitext(rep)[0] + ".raw.csv")  # tools/ncu_summary.sh exports this on the box
    if os.path.exists(csv_path):
        raw = open(csv_path).read()
    elif os.path.exists(path):
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        return None
    rows = list(csv.reader([l for l in raw.splitlines() if l.startswith('"')]))
    hdr, units = rows[0], rows[1]
    out = [f"# {tag}: {title}", "", "`ncu --set full --clock-control none --import-source on`, read with `--page raw --csv`.", ""]
    res = []
    for vals in rows[2:]:
        d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
        out.append(f"## `{d['Kernel Name'][1]}`  grid {d.get('Grid Size', ('', '?'))[1]} block {d.get('Block Size', ('', '?'))[1]}")
        out.append("")
        out.append("| metric | value | unit |")
        out.append("|---|---|---|")
        for k in KEYS:
            if k in d:
                out.append(f"| {k} | {d[k][1]} | {d[k][0]} |")
        out.append("")
        res.append(d)
    open(os.path.join(PROF, f"{tag}_{os.path.splitext(rep)[0]}.md"), "w").write("\n".join(out) + "\n")
    return res


def to_bytes(unit, val):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(val) * scale


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    launches(tag)
    res = ncu_full(tag, "prof_s1.ncu-rep", "dominant kernel (S1: 100000 instances x 20000 RK4 steps, log every 100)")
    bj = os.path.join(OUT, "bench.json")
    if os.path.exists(bj) and os.path.getsize(bj):
        open(os.path.join(PROF, f"{tag}_bench.json"), "w").write(open(bj).read())
    cj = os.path.join(OUT, "configs.jsonl")
    if os.path.exists(cj) and os.path.getsize(cj):
        open(os.path.join(PROF, f"{tag}_configs.jsonl"), "w").write(open(cj).read())
    if res:
        d = res[0]
        rd = to_bytes(*d["dram__bytes_read.sum"])
        wr = to_bytes(*d["dram__bytes_write.sum"])
        # the capture ran 20000 of the bench launch's 200000 steps per instance: DRAM traffic is the log
        # (proportional to steps) plus a per-launch constant; scale by the step ratio.
        scale = 10.0
        info = {
            "source": f"profiles/{tag}_prof_s1.md (ncu --set full, 100000 instances x 20000 steps)",
            "dram_bytes_captured_launch": rd + wr,
            "scale_to_bench_launch": scale,
            "dram_bytes_per_launch": (rd + wr) * scale,
        }
        try:
            # executed FP64 instruction mix of the dominant kernel, per layer-timestep
            f = lambda k: float(d[k][1].replace(",", ""))
            dfma = f("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed")
            dmul = f("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed")
            dadd = f("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed")
            all_per_cycle = f("sm__inst_executed.sum.per_cycle_elapsed") * f("smsp__thread_inst_executed_per_inst_executed.ratio")
            total = f("thread_inst_executed")
            layer_steps = 100000 * 20000 * 2
            per = total / layer_steps / all_per_cycle
            info.update({
                "thread_inst_per_layer_step": total / layer_steps,
                "dfma_per_layer_step": dfma * per,
                "dmul_per_layer_step": dmul * per,
                "dadd_per_layer_step": dadd * per,
                "fp64_thread_inst_per_layer_step": (dfma + dmul + dadd) * per,
                "executed_flop_per_layer_step": (2 * dfma + dmul + dadd) * per,
                "fp64_pipe_active_pct": f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            })
        except (KeyError, ValueError, ZeroDivisionError) as e:
            info["instruction_mix_error"] = repr(e)
        json.dump(info, open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
    for cfg, title in (("s3", "S3 p-bit: 1-layer STT, 300 K, Heun (first captured launch), 1e5 instances x 5000 steps"),
                       ("s4", "S4 CIMS: 3-layer Heun with SOT step drivers, reduced grid"),
                       ("s5s", "S5 SeriesStack of 64 MTJs, batch of stacks x 2000 steps"),
                       ("s5g", "S5 GroupInteraction 64 x 64 lattice (cooperative grid kernel)"),
                       ("mombg", "PIMM_batch: background Chebyshev-moment kernel of one chunk, running beside the time loop of the next")):
        if os.path.exists(os.path.join(OUT, f"prof_{cfg}.ncu-rep")):
            ncu_full(tag, f"prof_{cfg}.ncu-rep", title)
    print("wrote", sorted(os.listdir(PROF)))


if __name__ == "__main__":
    main()
