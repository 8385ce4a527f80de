"""Developer aid (GPU box): cycle time against the number of pipelined slices; checks results are slicing-independent."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from eepacc_mpc_casadi_matlab_b200 import api
cn = sys.argv[1] if len(sys.argv) > 1 else "ab"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
nst = int(sys.argv[3]) if len(sys.argv) > 3 else 871
gl = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1, 2, 4, 8, 16, 32]
p, routes, lead = bench.make_workload(cn, B, 0, nst)
eng = api.Engine(cn, p, routes, B=B)
eng.set_lead(lead)
ref = None
for G in gl:
    eng.set_groups(G)
    eng.run(nst, want_log=False)
    t0 = time.perf_counter()
    log, kpi = eng.run(nst, want_log=(G in (gl[0], gl[-1])))
    dt = time.perf_counter() - t0
    if log is not None:
        if ref is None:
            ref = log
        else:
            print("   log identical to first slicing:", np.array_equal(ref, log))
    print("groups %2d: %.3f s  %.0f QP/s  failed %d" % (eng.groups(), dt, B * nst / dt, int(eng.counters()[:, 1].sum())), flush=True)
