"""Summarise `ncu --set full` raw CSV pages into a small markdown table (development tool).
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv ; python tools_ncu_summary.py raw.csv"""
import csv, sys
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "tensor pipe inst %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__cycles_active.avg", "SMSP cycles active"), ("sm__cycles_elapsed.max", "SM cycles elapsed"),
]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
print(f"# {sys.argv[1]}")
for r in rows[2:]:
    print(f"\n## {r[col['Kernel Name']][:110]}\n")
    print("| metric | value |\n|---|---|")
    for k, label in KEYS:
        if k in col and r[col[k]] not in ("", "n/a"):
            print(f"| {label} | {r[col[k]]} {units[col[k]]} |")
    stalls = [(float(r[i]), h.split('stalled_')[1].split('_per_issue')[0]) for i, h in enumerate(hdr)
              if 'average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio') and r[i] not in ("", "n/a")]
    top = sorted(stalls, reverse=True)[:6]
    print("| top stalls (warps per issue) | " + ", ".join(f"{n} {v:.2f}" for v, n in top) + " |")
