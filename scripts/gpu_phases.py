"""Developer aid (GPU box): solver phase cycle counts (vehicle 0) over a few closed-loop steps."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from eepacc_mpc_casadi_matlab_b200 import api, settings as S
cn = sys.argv[1] if len(sys.argv) > 1 else "ab"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 148
nst = int(sys.argv[3]) if len(sys.argv) > 3 else 30
p, routes, lead = bench.make_workload(cn, B, 0, max(nst, 25))
eng = api.Engine(cn, p, routes, B=B)
eng.debug_trace()
names = ["load", "warmQR", "refacR", "auxsetup", "kkt", "Adx+ratio", "wschange", "refine+out",
         "k:usolveT", "k:QmulY", "k:Hmul", "k:QtmulZ", "k:Rsolves", "k:QmulZ", "k:QtmulY", "k:tsolveT",
         "r:Amul", "r:ratio", "r:update", "r:-",
         "w:classify", "w:plan", "w:stage", "w:QR", "w:Tfill", "w:formQ", "w:reverse", "w:-"]
tot = np.zeros(28)
for k in range(nst):
    out = eng.step(lead[k])
    ph = eng.debug_phases()
    if k in (0, 1, 5, nst - 1):
        print(k, "iter", out["iter"][0], {n_: int(v) for n_, v in zip(names, ph)}, "sum", int(ph[:8].sum()))
    if k > 0:
        tot += ph
print("mean over warm steps:", {n_: int(v / (nst - 1)) for n_, v in zip(names, tot)}, "sum", int(tot[:8].sum() / (nst - 1)))
