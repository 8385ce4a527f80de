"""Developer aid: dump the per-step QP operands K2 hands to K3 (structured block, g, bounds) and K3's answers
for a few closed-loop vehicles, so that the solver can be replayed off-line (tests/host_emu)."""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_workload
from eepacc_mpc_casadi_matlab_b200 import api
ctrl, B, ns = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
first = int(sys.argv[4]) if len(sys.argv) > 4 else 0
p, routes, lead = make_workload(ctrl, B, first, ns)
eng = api.Engine(ctrl, p, routes, B=B)
n, m = eng.n, eng.m
keys = ["sblk", "g", "lba", "uba", "x", "status", "iter", "cost"]
widths = [1304, n, m, m, n, 1, 1, 1]
acc = {k: [] for k in keys}
for k in range(ns):
    eng.step(lead[k])
    for i, (nm, w) in enumerate(zip(keys, widths)):
        acc[nm].append(eng.debug_fetch(i, B * w).reshape(B, w).copy())
out = {k: np.stack(v) for k, v in acc.items()}
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed(f"gpurun_out/qpdump_{ctrl}.npz", n=n, m=m, **out)
print(ctrl, "dumped", B, ns, "failed", int((out["status"] != 0).sum()), "mean iter", out["iter"].mean())
