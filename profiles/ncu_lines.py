#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line: python profiles/ncu_lines.py report.ncu-rep [top_n]
(runs `ncu -i ... --page source --csv --print-source cuda,sass` and aggregates the stall samples per line)."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, fname, out = None, "", []
    for r in rows:
        if r and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0].isdigit():
            out.append((fname, r))
    i_s = hdr.index("# Samples")
    i_inst = hdr.index("Instructions Executed")
    stall = [(j, h) for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[i_s]) for _, r in out if r[i_s].isdigit())
    print("total samples", tot)
    for f, r in sorted(out, key=lambda x: -int(x[1][i_s] or 0))[:top_n]:
        st = sorted([(int(r[j]), h[6:]) for j, h in stall if r[j].isdigit() and int(r[j])], reverse=True)[:3]
        print(f"{f}:{r[0]:>4} {int(r[i_s]):7d} {int(r[i_s]) / max(tot, 1):6.3f} inst={r[i_inst]:>9} {r[1].strip()[:80]}  {st}")


if __name__ == "__main__":
    main()
