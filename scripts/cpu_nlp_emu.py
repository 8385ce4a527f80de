"""CPU: the NLP benchmark solver (K7) under the host emulation of its device code (tests/host_emu/k7_emu.cpp) on the
reference's golden case and on perturbed full-cycle scenarios, each answer checked with the oracle's restatement of the
problem.  Writes the table DESIGN.md section 3 (K7) quotes:  python scripts/cpu_nlp_emu.py [n_scenarios] > profiles/r2_cpu_k7_emu.txt"""
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_nlp_host_emu as T                                                   # noqa: E402
from eepacc_mpc_casadi_matlab_b200 import nlp, settings as S, usecase as U      # noqa: E402
from oracle import oracle as O                                                   # noqa: E402

if not os.path.exists(T.EMU_LIB) or any(os.path.getmtime(f) > os.path.getmtime(T.EMU_LIB) for f in [T.EMU_SRC] + T.HDRS):
    os.makedirs(os.path.dirname(T.EMU_LIB), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", "/usr/local/cuda/include", "-o", T.EMU_LIB, T.EMU_SRC])
L = C.CDLL(T.EMU_LIB)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
lead = T.golden("lead_TO01_EAD.npz")["s_tv"]
N = 870
for tree in ("mc", "abo"):
    g = T.golden(f"nlp_{tree}.npz")
    p, r = S.default_params(tree), U.default_route()
    t = nlp.nlp_tables(r, p)
    t0 = time.time()
    z, st = T.emu_solve(L, p, r, t, lead, N)
    dt = time.time() - t0
    X = np.concatenate([z[:4][None], z[4:].reshape(N, 10)[:, 6:]])
    Uz = z[4:].reshape(N, 10)[:, :6]
    ev = O.nlp_eval(p, r, *T.tables_of(t), lead, z)
    print(f"golden {tree}: {nlp.STATUS[int(st[0])]}, {int(st[1])} iterations, {dt:.1f} s on one core (ipopt: {float(g['tSolve']):.0f} s), reported kkt {st[2]:.2e}, "
          f"oracle feasibility {O.nlp_violation(ev, z):.1e}, f {ev['f']:.10e}; vs savedNLPsol.mat: s {np.abs(X[:, 0] - g['s_opt']).max():.2e} m, "
          f"v {np.abs(X[:, 1] - g['v_opt']).max():.2e} m/s, Fm {np.abs(Uz[:, 0] - g['Fm_opt']).max():.2e} N")
p = S.default_params("mc")
rng = np.random.default_rng(1)
print("perturbed full-cycle scenarios (MC tree): mass x U[0.9, 1.1], lead speed x U[0.85, 1.15]")
for b in range(n):
    r = U.default_route()
    r.mass_scale = 1.0 if b == 0 else rng.uniform(0.9, 1.1)
    sc = 1.0 if b == 0 else rng.uniform(0.85, 1.15)
    s_tv = 6.0 + (lead - 6.0) * sc
    t = nlp.nlp_tables(r, p)
    t0 = time.time()
    z, st = T.emu_solve(L, p, r, t, s_tv, N)
    dt = time.time() - t0
    ev = O.nlp_eval(p, r, *T.tables_of(t), s_tv, z)
    s = np.r_[z[0], z[4:].reshape(N, 10)[:, 6]]
    corner = [int(k) for k in range(N + 1) if min(abs(s[k] - 994.4444444444445), abs(s[k] - 999.0), abs(s[k] - 1000.0)) < 1e-6]
    print(f"{b:2d} mass {r.mass_scale:.3f} lead {sc:.3f}: {nlp.STATUS[int(st[0])]:30s} {int(st[1]):3d} iterations, kkt {st[2]:.2e}, mu {st[5]:.1e}, "
          f"f {st[8]:.8e}, oracle feasibility {O.nlp_violation(ev, z):.1e}, {dt:.1f} s, stages within 1e-6 m of a table corner: {corner}")
