"""Summarise the per-call traces written with PDEB200_E2E_TRACE=<prefix> (pdeb200_act_step_host): per shard the mean
duration of each stage, and over all shards how busy the GPU kernels / the D2H copy engine were.
usage: PDEB200_E2E_TRACE=gpurun_out/tr python bench.py ...; python tools/e2e_timeline.py gpurun_out/tr"""
import glob, sys
import numpy as np

rows = []
for k, f in enumerate(sorted(glob.glob(sys.argv[1] + ".*.csv"))):
    a = np.loadtxt(f, delimiter=",", ndmin=2)
    if a.shape[0] < 20:
        continue
    a = a[len(a) // 2:]                      # steady state: second half (the first calls include warm-up legs)
    rows.append(a)
    h2d, kern, d2h = a[:, 1] - a[:, 0], a[:, 2] - a[:, 1], a[:, 3] - a[:, 2]
    call = a[:, 5] - a[:, 4]
    period = np.diff(a[:, 0]).mean()
    print("shard %d: calls %d  period %.1f us | H2D %.1f  kernels %.1f  D2H %.1f  (device chain %.1f) | host call %.1f us, gap between calls %.1f us" % (
        k, len(a), 1e3 * period, 1e3 * h2d.mean(), 1e3 * kern.mean(), 1e3 * d2h.mean(), 1e3 * (a[:, 3] - a[:, 0]).mean(),
        1e3 * call.mean(), 1e3 * (a[1:, 4] - a[:-1, 5]).mean()))
if rows:
    t0 = max(r[0, 0] for r in rows); t1 = min(r[-1, 3] for r in rows)
    def busy(lo, hi):
        iv = sorted((max(r[i, lo], t0), min(r[i, hi], t1)) for r in rows for i in range(len(r)) if r[i, hi] > t0 and r[i, lo] < t1)
        tot, cur_s, cur_e = 0.0, None, None
        for s, e in iv:
            if cur_e is None or s > cur_e:
                if cur_e is not None: tot += cur_e - cur_s
                cur_s, cur_e = s, e
            else:
                cur_e = max(cur_e, e)
        if cur_e is not None: tot += cur_e - cur_s
        return tot / (t1 - t0)
    print("window %.2f ms: some kernel stage active %.0f %% of the time, some D2H stage %.0f %%, some H2D stage %.0f %%" % (
        t1 - t0, 100 * busy(1, 2), 100 * busy(2, 3), 100 * busy(0, 1)))
