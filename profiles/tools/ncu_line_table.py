#!/usr/bin/env python
"""Per-source-line instruction / sample / shared-wavefront totals of a kernel from an .ncu-rep captured with
--import-source on (needs -lineinfo):  python profiles/tools/ncu_line_table.py rep.ncu-rep file.cu [lo:hi:name ...]"""
import csv
import subprocess
import sys


def main():
    rep, fname = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    hdr, cur, out = None, None, []
    for r in csv.reader(txt.splitlines()):
        if r and r[0] == "File Path":
            cur = r[1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0] and cur and cur.endswith(fname):
            d = {}
            for k, v in zip(hdr, r):
                d.setdefault(k, v)
            try:
                out.append((int(r[0]), int(d["Instructions Executed"]), int(d["# Samples"]),
                            int(d["L1 Wavefronts Shared"] or 0), int(d.get("stall_barrier") or 0), r[1][:80]))
            except ValueError:
                pass
    ti, ts, tw = (sum(o[i] for o in out) for i in (1, 2, 3))
    print(f"total: inst {ti} samples {ts} shared wavefronts {tw}")
    for spec in sys.argv[3:]:
        lo, hi, name = spec.split(":")
        s = [o for o in out if int(lo) <= o[0] <= int(hi)]
        print(f"{name:10s} lines {lo}-{hi}: inst {sum(o[1] for o in s) / max(ti, 1):.3f} samples "
              f"{sum(o[2] for o in s) / max(ts, 1):.3f} (barrier {sum(o[4] for o in s) / max(ts, 1):.3f}) "
              f"wavefronts {sum(o[3] for o in s) / max(tw, 1):.3f}")
    for o in sorted(out, key=lambda o: -o[1])[:40]:
        print(f"{o[0]:5d} inst {o[1]:9d} samples {o[2]:5d} wf {o[3]:9d}  {o[5]}")


if __name__ == "__main__":
    main()
