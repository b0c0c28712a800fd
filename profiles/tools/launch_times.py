"""Per-kernel launch durations (ms) from an `ncu --metrics gpu__time_duration.sum --csv` log; optional name filter."""
import collections
import csv
import sys

pat = sys.argv[2:] or [""]
hdr, agg = None, collections.OrderedDict()
for r in csv.reader(open(sys.argv[1])):
    if len(r) <= 5:
        continue
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    name = d["Kernel Name"][:64]
    if not any(p in name for p in pat):
        continue
    try:
        v = float(d["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    v *= {"us": 1e-3, "ns": 1e-6, "ms": 1.0, "s": 1e3}.get(d["Metric Unit"], 1.0)
    agg.setdefault(name, []).append(v)
for k, v in agg.items():
    print(f"{k:66s} n={len(v):3d} total={sum(v):9.3f} ms :", " ".join(f"{x:.2f}" for x in v[:24]))
