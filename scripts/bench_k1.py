"""BASELINE config 5: batched f / F / dF evaluation over 10^6 points (device-resident), CUDA events,
HBM roofline (algorithmic bytes/point: f 120, F 120, F+Jacobian 520) and oracle spot-check."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eepacc_mpc_casadi_matlab_b200 import api, settings as S, usecase as U

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "nlp_mc.npz"))
m = len(g["Fm_opt"])
rng = np.random.default_rng(5)
base_x = np.c_[g["s_opt"][:m], g["v_opt"][:m], g["theta_opt"][:m], np.r_[g["j_opt"], 0.0][:m]]
base_u = np.c_[g["Fm_opt"], g["Fb_opt"], np.zeros((m, 4))]
rep = (n + m - 1) // m
x = np.tile(base_x, (rep, 1))[:n] * (1 + 1e-3 * rng.standard_normal((n, 4)))
u = np.tile(base_u, (rep, 1))[:n] * (1 + 1e-3 * rng.standard_normal((n, 6)))
p = S.default_params("mc")
eng = api.Engine("fb", p, U.default_route(), B=1)
xd, ud = torch.from_numpy(x).cuda(), torch.from_numpy(u).cuda()
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {}
for name, fn, nbytes in (("f", lambda: eng.f(xd, ud), 120), ("F", lambda: eng.F(xd, ud), 120),
                         ("F_jac", lambda: eng.F_jac(xd, ud), 520)):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    gbs = n * nbytes / (ms * 1e-3) / 1e9
    out[name] = {"ms": ms, "points_per_s": n / (ms * 1e-3), "GB/s": gbs, "frac_hbm": gbs / peaks["hbm_gbs"]}
# spot check against the oracle
from oracle import oracle as O
L = O.lib(); dp = C.POINTER(C.c_double)
xf, qf, J = [t.cpu().numpy() for t in eng.F_jac(xd, ud)]
worst = 0.0
for b in range(0, n, max(1, n // 200)):
    a, q, Jo = np.zeros(4), np.zeros(1), np.zeros(50)
    L.orc_F_jac(C.byref(p), 1.0, x[b].ctypes.data_as(dp), u[b].ctypes.data_as(dp), a.ctypes.data_as(dp), q.ctypes.data_as(dp), Jo.ctypes.data_as(dp))
    Jo = Jo.reshape(10, 5).T
    worst = max(worst, np.max(np.abs(xf[b] - a) / (1 + np.abs(a))), abs(qf[b] - q[0]) / (1 + abs(q[0])),
                np.max(np.abs(J[b] - Jo) / (1e-9 + np.abs(Jo).max(axis=1, keepdims=True))))
out["n_points"] = n
out["max_rel_err_vs_oracle"] = float(worst)
print(json.dumps(out))
