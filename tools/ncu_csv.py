"""Summarise an `ncu --csv` log: one line per launch with the requested metrics (rows are one metric per line)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = None
launches = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r and "Metric Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    k = (d["ID"], d["Kernel Name"][:48], d.get("Grid Size", ""), d.get("Block Size", ""))
    launches.setdefault(k, {})[d["Metric Name"]] = (d["Metric Value"], d["Metric Unit"])
for k, m in launches.items():
    print(k[0], k[1], k[2], k[3], " ".join(f"{n.split('.')[0].replace('smsp__','').replace('gpu__','')}={v}{u}" for n, (v, u) in m.items()))
