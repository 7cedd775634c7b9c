"""Top SASS instructions by stall samples from `ncu --page source --csv` (SASS view), with executed counts."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
data = []
for i, r in enumerate(rows[2:]):
    try:
        data.append((i, r[1].strip(), int(r[2]), int(r[5]), float(r[8]) if r[8] not in ("", "-") else 0.0))
    except (ValueError, IndexError):
        pass
tot = sum(d[2] for d in data)
toti = sum(d[3] for d in data)
print("total samples", tot, "total warp-instr", toti)
for d in sorted(data, key=lambda d: -d[2])[:top]:
    print(f"{d[0]:5d} {100*d[2]/tot:5.1f}%  exec={d[3]:10d} thr={d[4]:4.1f}  {d[1]}")
