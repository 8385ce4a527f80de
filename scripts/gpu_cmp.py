"""Developer aid (GPU box): step-by-step CUDA vs oracle for one perturbed bench scenario."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from eepacc_mpc_casadi_matlab_b200 import api, settings as S
from oracle import oracle as O
np.set_printoptions(linewidth=220, precision=9)
cn, b, nst = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
p, routes, lead = bench.make_workload(cn, b + 1, 0, max(nst, 2))
eng = api.Engine(cn, p, routes[b], B=1)
V = O.Vehicle(S.CTRL_NAMES[cn], p, routes[b])
for k in range(nst):
    out = eng.step(lead[k, b:b + 1])
    o = V.step(*lead[k, b])
    d = out["z1"][0] - o[:8]
    print(k, "lead", lead[k, b], "dev", np.abs(d).max(), "sumdev", (out["z1"][0][2] + out["z1"][0][3]) - (o[2] + o[3]))
    print("   gpu", out["z1"][0], out["status"], out["iter"])
    print("   orc", o[:8], o[9], o[10], "cost", o[11])
