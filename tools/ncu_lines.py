#!/usr/bin/env python3
"""Warp-state samples of one launch of an .ncu-rep aggregated per CUDA-C source line (needs -lineinfo and
--import-source on).  usage: tools/ncu_lines.py rep launch_index [top_n]"""
import collections
import csv
import io
import subprocess
import sys


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


rep, skip = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(skip),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# find header rows
views = [i for i, r in enumerate(rows) if "# Samples" in r]
for vi, start in enumerate(views):
    h = rows[start]
    sx = {x: i for i, x in enumerate(h)}
    end = views[vi + 1] - 1 if vi + 1 < len(views) else len(rows)
    data = [r for r in rows[start + 1:end] if len(r) >= len(h)]
    tot = sum(num(r[sx["# Samples"]]) for r in data) or 1
    if "Address" in sx and vi == 0 and len(views) > 1:
        continue        # SASS view first; the CUDA-C view follows
    print("view %d: %d rows, %d samples" % (vi, len(data), tot))
    key = "Source"
    agg = collections.Counter()
    ins = collections.Counter()
    for r in data:
        agg[r[sx[key]].strip()[:110]] += num(r[sx["# Samples"]])
        ins[r[sx[key]].strip()[:110]] += num(r[sx["Instructions Executed"]])
    ti = sum(ins.values()) or 1
    for s, v in agg.most_common(top):
        print("%5.1f%% samples %5.1f%% inst | %s" % (100.0 * v / tot, 100.0 * ins[s] / ti, s))
