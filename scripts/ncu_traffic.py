#!/usr/bin/env python
"""Read a `ncu --set full` capture of k3r_qp_kernel and (1) print the judged counters, (2) store the DRAM bytes per
vehicle in profiles/k3r_traffic.json (bench.py's roofline.traffic).  usage: ncu_traffic.py <rep> <ctrl> <vehicles>"""
import csv
import io
import json
import os
import subprocess
import sys

rep, ctrl, veh = sys.argv[1], sys.argv[2], int(sys.argv[3])
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warp_latency_issue_stalled_barrier.pct", "sm__cycles_active.avg", "sm__cycles_elapsed.max"]
KEYS += ["sm__inst_executed.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
         "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
         "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
         "sm__sass_inst_executed_op_shared_ld.sum", "sm__sass_inst_executed_op_shared_st.sum", "smsp__inst_executed.sum",
         "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
KEYS += sorted(k for k in hdr if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")
               and k not in KEYS)
out = []
tot = []
fp64 = []
for d in data:
    rec = dict(zip(hdr, d))
    out.append("--- launch ---")
    for k in KEYS:
        if k in rec:
            out.append(f"{k:75s} {rec[k]:>18s} {units[hdr.index(k)]}")
    if "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active" in rec:
        fp64.append(float(rec["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"].replace(",", "")) / 100.0)

    def num(k):
        v = float(rec[k].replace(",", ""))
        u = units[hdr.index(k)].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    tot.append(num("dram__bytes_read.sum") + num("dram__bytes_write.sum"))
print("\n".join(out))
per = sum(tot) / len(tot) / veh
print(f"DRAM bytes per launch (mean of {len(tot)}): {sum(tot) / len(tot):.0f}  -> per vehicle {per:.1f}")
tj_path = os.path.join(ROOT, "profiles", "k3r_traffic.json")
tj = json.load(open(tj_path))
tj[ctrl] = {"dram_bytes_per_vehicle": per, "vehicles_per_launch": veh,
            "fp64_pipe_active_frac": (sum(fp64) / len(fp64)) if fp64 else None,
            "capture": f"profiles/{os.path.basename(rep).replace('.ncu-rep', '')}_summary.txt: ncu --set full, sm__pipe_fp64_cycles_active "
                       f"(share of active cycles) of {len(tot)} k3r_qp_kernel launches of {veh} vehicles",
            "source": f"profiles/{os.path.basename(rep).replace('.ncu-rep', '')}_summary.txt (dram__bytes_read.sum + "
                      f"dram__bytes_write.sum, mean of {len(tot)} k3r_qp_kernel launches)"}
json.dump(tj, open(tj_path, "w"), indent=1)
