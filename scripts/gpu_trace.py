"""Developer aid (GPU box): per-iteration solver trace of the first step, CUDA vs oracle."""
import sys, os, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eepacc_mpc_casadi_matlab_b200 import api, settings as S, usecase as U

cn = sys.argv[1] if len(sys.argv) > 1 else "tv"
step = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lead = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "lead_TO01_EAD.npz"))
L = np.stack([lead["s_tv"], lead["v_tv"]], 1)
p = S.default_params("mc"); r = U.default_route()
eng = api.Engine(cn, p, r, B=1)
eng.debug_trace()
for k in range(step + 1):
    out = eng.step(L[k:k + 1])
    tr = eng.debug_trace()
mt = eng.n + eng.m
print("gpu status", out["status"], "iter", out["iter"], "z1", out["z1"][0][:4])
for i, (tau, key, nA, ex) in enumerate(tr[:int(sys.argv[3]) if len(sys.argv) > 3 else 150]):
    key = int(key)
    if key < 2 * mt:
        print("it %d tau %.3e blk %d side %d nA %d ex %d" % (i + 1, tau, key >> 1, 1 if key & 1 else -1, nA, ex))
    else:
        print("it %d tau %.3e dualpos %d nA %d ex %d" % (i + 1, tau, key - 2 * mt, nA, ex))
