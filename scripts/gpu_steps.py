"""Developer aid (GPU box): per-step wall time, working-set-change statistics, deferred vehicles and failures
over a closed-loop cycle.  usage: gpu_steps.py ctrl B n_steps"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from eepacc_mpc_casadi_matlab_b200 import api
cn = sys.argv[1] if len(sys.argv) > 1 else "ab"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
nst = int(sys.argv[3]) if len(sys.argv) > 3 else 200
p, routes, lead = bench.make_workload(cn, B, 0, nst)
eng = api.Engine(cn, p, routes, B=B)
rows = []
for k in range(nst):
    t0 = time.perf_counter()
    out = eng.step(lead[k])
    dt = (time.perf_counter() - t0) * 1e3
    it = out["iter"]; st = out["status"]
    dc = eng.debug_fetch(8, 2).view(np.int32)[0]
    rows.append((k, dt, it.mean(), np.percentile(it, 99), it.max(), int(dc), int((st != 0).sum())))
    if (st != 0).any():
        print("  step", k, "failed vehicles", np.nonzero(st != 0)[0][:10], "status", st[st != 0][:10])
r = np.array(rows)
print("step    ms   it_mean it_p99 it_max deferred failed")
order = np.argsort(-r[:, 1])
for i in list(range(0, min(nst, 12))) + [int(j) for j in order[:25]]:
    print("%4d %7.2f %6.2f %6.0f %6.0f %6d %4d" % tuple(r[i]))
print("total ms %.1f  median %.2f  mean %.2f  deferred total %d (%.3f%% of QPs)" % (
    r[:, 1].sum(), np.median(r[:, 1]), r[:, 1].mean(), r[:, 5].sum(), 100 * r[:, 5].sum() / (B * nst)))
# time by bucket of it_max
for lo, hi in [(0, 10), (10, 20), (20, 40), (40, 80), (80, 1e9)]:
    sel = (r[:, 4] >= lo) & (r[:, 4] < hi)
    print("it_max in [%g,%g): %d steps, %.1f ms" % (lo, hi, sel.sum(), r[sel, 1].sum()))
