"""
Warp-stall breakdown per kernel from `ncu -i <rep> --page raw --csv`
(smsp__average_warps_issue_stalled_*_per_issue_active: warps waiting per issued instruction).

    python tools/ncu_stalls.py gpurun_out/prof_raw.csv
"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = rows[0]
    name_i = hdr.index("Kernel Name")
    idx = [i for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    for r in rows[2:]:
        vals = []
        for i in idx:
            try:
                vals.append((float(r[i]), hdr[i][len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
        vals.sort(reverse=True)
        print(r[name_i][:60].replace("void ", ""), " ".join(f"{n}={v:.2f}" for v, n in vals[:6]))


if __name__ == "__main__":
    main()
