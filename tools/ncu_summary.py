"""Summarise ncu raw-page CSVs (one kernel per file) into a small table. usage: ncu_summary.py file.raw.csv ..."""
import csv
import sys

WANT = [
    ("Kernel Name", "kernel"), ("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_nominal"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"), ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"), ("launch__grid_size", "grid"),
    ("launch__block_size", "block"), ("smsp__inst_executed.sum", "warp_insts"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_insts"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
]
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print(f"== {path}")
        for key, label in WANT:
            for i, h in enumerate(hdr):
                if h == key:
                    print(f"  {label:22s} {vals[i]} {units[i]}")
