#!/usr/bin/env python3
"""Warp-state samples of one launch of an .ncu-rep aggregated per CUDA-C source line (needs -lineinfo and
--import-source on).  usage: tools/ncu_lines.py rep launch_index [top_n]
A SASS instruction inlined from several files is listed under each of them (the shares of different files overlap)."""
import collections
import csv
import io
import os
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
agg, ins, text, stalls = collections.Counter(), collections.Counter(), {}, collections.defaultdict(collections.Counter)
fname, hdr = "?", None
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        fname = os.path.basename(r[1]); hdr = None
        continue
    if "# Samples" in r:
        hdr = r
        si, ii = r.index("# Samples"), r.index("Instructions Executed")
        st_cols = [(i, x) for i, x in enumerate(r) if x.startswith("stall_") and "Not Issued" not in x]
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if not r[0].strip():            # SASS detail rows under a source line (the line's own row carries the aggregate)
        continue
    key = (fname, r[0])
    agg[key] += num(r[si]); ins[key] += num(r[ii]); text[key] = r[1].strip()[:90]
    for i, x in st_cols:
        stalls[key][x] += num(r[i])
# totals from the plain SASS view (in the cuda,sass view an instruction is repeated under every source line it belongs to)
out2 = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(skip),
                       "--launch-count", "1"], capture_output=True, text=True).stdout
tot = ti = 0
h2 = None
for r in csv.reader(io.StringIO(out2)):
    if "# Samples" in r:
        h2 = r; s2, i2 = r.index("# Samples"), r.index("Instructions Executed")
        continue
    if h2 is not None and len(r) >= len(h2):
        tot += num(r[s2]); ti += num(r[i2])
tot, ti = tot or 1, ti or 1
print("samples %d, warp instructions %d" % (tot, ti))
for key, v in agg.most_common(top):
    s2 = ", ".join("%s %.0f%%" % (k[6:], 100.0 * c / max(v, 1)) for k, c in stalls[key].most_common(2))
    print("%5.1f%% smp %5.1f%% ins | %-14s:%-4s | %-90s | %s" % (100.0 * v / tot, 100.0 * ins[key] / ti, key[0], key[1], text[key], s2))
