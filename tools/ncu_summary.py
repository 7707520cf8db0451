#!/usr/bin/env python
"""Summarises an .ncu-rep (ncu --set full) into a small CSV of the metrics DESIGN.md quotes.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_trace_full.csv [workload-name]
With a workload name (bench.py's SCENE) it also writes profiles/ncu_traffic.json: the DRAM bytes of the incoherent closest-hit
launch, which bench.py reports as roofline.traffic.  The capture holds four consecutive trace launches of bench.py's step
(coherent closest, incoherent closest, coherent occluded, incoherent occluded, cyclically): the incoherent closest-hit launch
is the second of the two adjacent trace_kernel<0,...> launches."""
import csv
import json
import os
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "l1tex__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    if rep.endswith(".csv"):   # already exported on the GPU box: ncu -i x.ncu-rep --page raw --csv (tools/gpu_profile_r2.sh)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [(w, hdr.index(w)) for w in WANT if w in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(data))])
        for name, i in cols:
            w.writerow([name, units[i]] + [r[i] for r in data])
    print(open(out).read())
    if len(sys.argv) > 3 and len(data) == 4:
        ki = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
        names = [r[ki] for r in data]
        closest = ["trace_kernel<0" in x or "trace_flat_kernel<0" in x for x in names]
        pick = [i for i in range(4) if closest[i] and closest[(i - 1) % 4]]
        if len(pick) == 1:
            i = pick[0]
            val = lambda m: float(data[i][hdr.index(m)].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[hdr.index(m)]]
            sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            import bench
            t = {"workload": sys.argv[3], "kernel": names[i], "kernel_source_hash": bench.kernel_source_hash(), "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                 "dram_read": val("dram__bytes_read.sum"), "dram_write": val("dram__bytes_write.sum"), "source": os.path.basename(out)}
            json.dump(t, open(os.path.join(os.path.dirname(out), "ncu_traffic.json"), "w"), indent=1)
            print("traffic", t)


if __name__ == "__main__":
    main()
