"""Developer aid (GPU box, library built with EEP_K3_TIMERS=1): clock64() cycles of the reduced-space
QP kernel by phase, vehicle 0, averaged over warm closed-loop steps."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from eepacc_mpc_casadi_matlab_b200 import api
cn = sys.argv[1] if len(sys.argv) > 1 else "ab"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 148
nst = int(sys.argv[3]) if len(sys.argv) > 3 else 60
veh0 = int(sys.argv[4]) if len(sys.argv) > 4 else 0
p, routes, lead = bench.make_workload(cn, B, veh0, max(nst, 25))
eng = api.Engine(cn, p, routes, B=B)
eng.debug_trace()
names = ["load", "w:classify", "w:stage", "w:QR", "w:formQ", "refacR", "auxsetup", "k:range", "k:Hmul1", "k:null",
         "k:Hmul2", "k:dualG", "k:dualX", "Amul", "ratio+step", "ws:add", "ws:remove", "refine", "out"]
tot = np.zeros(28); cnt = 0; its = 0
for k in range(nst):
    out = eng.step(lead[k])
    ph = eng.debug_phases()
    if k >= 25:
        tot += ph; cnt += 1; its += out["iter"][0]
print(cn, "mean iter", its / cnt)
for n_, v in zip(names, tot / cnt):
    print(f"  {n_:12s} {int(v):8d}")
print("  sum", int(tot[:19].sum() / cnt))
