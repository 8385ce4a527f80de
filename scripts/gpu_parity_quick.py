"""Developer aid (GPU box): quick parity of the library selected by EEPACC_LIB against the oracle on the bench batch."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench
from eepacc_mpc_casadi_matlab_b200 import api, settings as S
import test_gpu_fb_headline as T
ctrl = sys.argv[1] if len(sys.argv) > 1 else "fb"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ns = int(sys.argv[3]) if len(sys.argv) > 3 else 200
p, routes, lead = bench.make_workload(ctrl, B, 0, ns)
eng = api.Engine(ctrl, p, routes, B=B)
eng.set_lead(lead)
log, kpi = eng.run(ns)
eng.close()
if ctrl == "fb":
    ref = T.oracle_logs(B, ns)
else:
    from oracle import oracle as O
    ref = np.stack([O.run(S.CTRL_NAMES[ctrl], p, routes[b], lead[:, b, :])[0] for b in range(B)], 1)
print(os.environ.get("EEPACC_LIB", "default lib"), ctrl, T.compare_logs(log, ref))
