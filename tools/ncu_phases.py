"""Share of the warp-stall samples between consecutive barriers of one kernel
(`ncu -i rep --page source --csv --kernel-name regex:NAME > src.csv; python tools/ncu_phases.py src.csv`):
for a multi-phase kernel this names the phase where the time goes."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
h = rows[hdr]
si, sm, ie = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
seen, data = set(), []
for r in rows[hdr + 1:]:
    if r[0] in seen:
        continue
    seen.add(r[0])
    try:
        data.append((float(r[sm]), r[si].strip(), float(r[ie] or 0), [float(r[c] or 0) for c in stall_cols]))
    except ValueError:
        pass
tot = sum(d[0] for d in data)
acc, ninstr, stalls, first = 0.0, 0.0, [0.0] * len(stall_cols), 0
for k, (v, s, e, st) in enumerate(data):
    acc += v
    ninstr += e
    stalls = [a + b for a, b in zip(stalls, st)]
    if s.startswith(("BAR", "UCGABAR_WAIT", "EXIT")) or k == len(data) - 1:
        if acc / tot > 0.005:
            top = sorted(zip(stalls, [h[c] for c in stall_cols]), reverse=True)[:3]
            print(f"{acc / tot * 100:5.1f}%  sass {first:5d}..{k:5d}  warp-instr {ninstr:9.0f}  ends at {s[:28]:28s} "
                  + " ".join(f"{n}={v / max(acc, 1) * 100:.0f}%" for v, n in top))
        acc, ninstr, stalls, first = 0.0, 0.0, [0.0] * len(stall_cols), k + 1
