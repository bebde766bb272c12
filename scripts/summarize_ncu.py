"""Summarise an .ncu-rep (ncu --set full) into a small markdown table for profiles/.
Usage: python scripts/summarize_ncu.py gpurun_out/x.ncu-rep > profiles/x.md"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
WANT = [
    ("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs/thread"),
    ("gpu__time_duration.sum", "duration"), ("sm__cycles_elapsed.avg.per_second", "SM clock"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe fma %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe alu %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe xu %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe lsu %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 sector hit %"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 global-load sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "L1 global-load requests"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall dispatch / issue"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle / issue"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving / issue"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction / issue"),
    ("smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio", "stall imc_miss / issue"),
    ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall sleeping / issue"),
    ("smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio", "stall tex_throttle / issue"),
    ("smsp__average_warps_issue_stalled_drain_per_issue_active.ratio", "stall drain / issue"),
    ("smsp__average_warps_issue_stalled_selected_per_issue_active.ratio", "selected / issue"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data-stage wavefronts %"),
    ("smsp__inst_executed_op_local_ld.sum", "local loads (spill)"),
    ("smsp__inst_executed_op_local_st.sum", "local stores (spill)"),
]
print(f"# ncu --set full --clock-control none: `{rep.split('/')[-1]}`\n")
print("| metric | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |")
print("|---|" + "---|" * len(data))
for key, label in WANT:
    if key not in idx:
        continue
    u = units[idx[key]]
    vals = []
    for r in data:
        v = r[idx[key]]
        if key == "Kernel Name":
            v = v.split("(")[0].replace("void ", "")
        else:
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
        vals.append(v)
    print(f"| {label}{' [' + u + ']' if u else ''} | " + " | ".join(vals) + " |")
