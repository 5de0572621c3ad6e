#!/usr/bin/env python
"""Stall samples of an .ncu-rep (captured with --import-source on, built with -lineinfo) summed per CUDA source line.

    python profiles/ncu_lines.py gpurun_out/<name>.ncu-rep [top]
"""
import collections, csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, agg, text = None, None, collections.Counter(), {}
for r in csv.reader(io.StringIO(out)):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r; ix = hdr.index("# Samples")
    elif hdr and len(r) == len(hdr) and r[0]:
        try:
            agg[(cur, int(r[0]))] += int(r[ix]); text[(cur, int(r[0]))] = r[1].strip()
        except ValueError:
            pass
tot = sum(agg.values())
print("samples", tot)
for (f, l), n in agg.most_common(top):
    print(f"{f[:20]:20s} {l:5d} {n:7d} {100 * n / tot:5.1f}%  {text[(f, l)][:110]}")
