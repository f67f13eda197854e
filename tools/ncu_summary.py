"""
Reduce `ncu -i <rep> --page raw --csv` to the columns the roofline discussion uses.

    python tools/ncu_summary.py gpurun_out/prof_all_r01_raw.csv > profiles/r01_ncu_all_kernels.csv
"""
import csv
import re
import sys

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp_insts"),
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = csv.writer(sys.stdout)
    out.writerow(["kernel"] + [f"{n} [{units[idx[k]]}]" if k in idx and units[idx[k]] else n for k, n in COLS])
    for r in data:
        name = r[idx["Kernel Name"]]
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", name))
        out.writerow([name] + [r[idx[k]] if k in idx else "" for k, _ in COLS])


if __name__ == "__main__":
    main()
