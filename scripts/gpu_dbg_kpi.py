import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from eepacc_mpc_casadi_matlab_b200 import api, settings as S, usecase as U
from oracle import oracle as O
np.set_printoptions(linewidth=250, precision=6, suppress=False)
p = S.default_params("mc")
d = np.load(os.path.join(bench.GOLDEN, "lead_TO01_EAD.npz"))
y = np.r_[0.0, np.diff(d["s_tv"]) / p.Ts][:871]
for B in (1, 3):
    routes, lead = U.perturbed_scenarios(B, y, n_steps=40)
    eng = api.Engine("ab", p, routes, B=B)
    eng.set_lead(lead)
    log, kpi = eng.run(40)
    print("B", B, "groups", eng.groups(), "fused", eng.fused())
    eng.close()
    for b in range(B):
        ref, kref = O.run(S.AB, p, routes[b], lead[:, b, :])
        print(" veh", b, "state err", np.abs(log[:, b, :2] - ref[:, :2]).max(), "status", log[:3, b, 9], ref[:3, 9], "iter", log[:3, b, 10], ref[:3, 10])
        print("  dev", kpi[b]); print("  orc", kref)
# fused test reproduction
B, ns = 96, 30
routes, lead = U.perturbed_scenarios(B, y, n_steps=ns)
for fused in (True, False):
    eng = api.Engine("ab", p, routes, B=B)
    eng.set_fused(fused); eng.set_lead(lead)
    log, kpi = eng.run(ns)
    print("fused", fused, "groups", eng.groups(), "status first 4 steps veh0..5:\n", log[:4, :6, 9], "\n iter:\n", log[:4, :6, 10])
    eng.close()
ref, _ = O.run(S.AB, p, routes[0], lead[:, 0, :])
print("oracle veh0 status", ref[:4, 9], "iter", ref[:4, 10])
