#!/usr/bin/env python
"""One line per kernel launch from an `ncu --page raw --csv` export: the counters the roofline discussion in DESIGN.md uses."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("us", "gpu__time_duration.sum"), ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size"), ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("fp64%", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        ("lsu%", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"), ("l2%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("gld_Msect", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"),
        ("l1hit%", "l1tex__t_sector_hit_rate.pct"), ("l2rd_Msect", "lts__t_sectors_srcunit_tex_op_read.sum"), ("dramR_MB", "dram__bytes_read.sum"), ("dramW_MB", "dram__bytes_write.sum"),
        ("loc_ld", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum"), ("loc_st", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"),
        ("st_long", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"), ("st_bar", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
        ("st_short", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"), ("st_lg", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
        ("st_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"), ("st_math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio")]
print("kernel".ljust(34) + "".join(n.rjust(10) for n, _ in cols))
def num(r, key):
    if key not in idx: return float("nan")
    v = r[idx[key]].replace(",", "")
    try: x = float(v)
    except ValueError: return float("nan")
    u = units[idx[key]]
    if key.startswith("dram__bytes"):
        x *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
    if "sectors" in key and key != "l1tex__t_sector_hit_rate.pct" and ("loc" not in key.replace("local", "loc") or True) and "local" not in key: x *= 1e-6
    if key == "gpu__time_duration.sum" and u == "ns": x *= 1e-3
    if key == "gpu__time_duration.sum" and u == "ms": x *= 1e3
    return x
for r in data:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
    print(name[:33].ljust(34) + "".join(f"{num(r, k):10.2f}" for _, k in cols))
