#!/usr/bin/env python3
"""Summarise an .ncu-rep (one kernel launch) into markdown for profiles/.

usage: tools/ncu_summary.py gpurun_out/foo.ncu-rep "title" [launch index] >> profiles/foo.md
Needs `ncu` on PATH (reading a report needs no GPU)."""
import collections
import csv
import io
import subprocess
import sys

RAW_KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), blocks"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), blocks"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe % of peak (active)"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe % of peak (active)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__cycles_elapsed.avg", "SM cycles elapsed"),
]


SKIP = 0


def ncu(rep, page):
    cmd = ["ncu", "-i", rep, "--page", page, "--csv", "--launch-skip", str(SKIP), "--launch-count", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    global SKIP
    rep, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
    SKIP = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # which launch of a multi-launch report
    raw = ncu(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[-1]
    ix = {h: i for i, h in enumerate(hdr)}
    print("## %s\n" % title)
    print("kernel: `%s`\n" % vals[ix["Kernel Name"]][:120])
    print("| metric | value |\n|---|---|")
    for k, name in RAW_KEYS:
        if k in ix:
            print("| %s (`%s`) | %s %s |" % (name, k, vals[ix[k]], units[ix[k]]))
    src = ncu(rep, "source")
    if len(src) > 3:
        h = src[1]
        sx = {x: i for i, x in enumerate(h)}
        data = []
        for r in src[2:]:
            if len(r) < len(h):
                continue
            if r[sx["# Samples"]] == "# Samples":      # second view (CUDA-C source) of the same launch follows
                break
            data.append(r)
        stall_cols = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
        tot = sum(int(r[sx["# Samples"]] or 0) for r in data) or 1
        st = collections.Counter()
        ops = collections.Counter()
        for r in data:
            for c in stall_cols:
                st[c] += int(r[sx[c]] or 0)
            s = r[sx["Source"]].strip()
            op = (s.split()[1] if s.startswith("@") else s.split()[0]).split(".")[0]
            ops[op] += int(r[sx["Instructions Executed"]] or 0)
        ti = sum(ops.values()) or 1
        print("\nSASS lines: %d; warp-state samples: %d\n" % (len(data), tot))
        print("| stall reason | % of samples |\n|---|---|")
        for c, v in st.most_common(8):
            print("| %s | %.1f |" % (c, 100.0 * v / tot))
        print("\n| opcode | % of executed warp instructions |\n|---|---|")
        for c, v in ops.most_common(10):
            print("| %s | %.1f |" % (c, 100.0 * v / ti))
    print()


if __name__ == "__main__":
    main()
