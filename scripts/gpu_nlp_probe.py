"""GPU probe of the batched NLP path (K7): a short-horizon batch against the host emulation of the same device code,
then the full golden cycle against savedNLPsol.mat.  No torch import (start-up time).  Output: stdout."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eepacc_mpc_casadi_matlab_b200 import nlp, settings as S, usecase as U     # noqa: E402

t00 = time.time()
lead = np.load(os.path.join(ROOT, "tests", "golden", "lead_TO01_EAD.npz"))["s_tv"]
p, r = S.default_params("mc"), U.default_route()


def batch(B, N):
    sc = 1.0 + 0.1 * (np.arange(B) % 7) / 7.0
    return np.stack([6.0 + (lead[:N + 1] - 6.0) * s for s in sc])


def run(B, N, tag):
    s_tv = batch(B, N)
    t0 = time.time()
    S_ = nlp.NlpSolver(p, r, N, B=B)
    t1 = time.time()
    sol = S_.solve(s_tv)
    t2 = time.time()
    S_.close()
    print(f"{tag}: B {B} N {N} setup {t1 - t0:.2f} s solve wall {t2 - t1:.2f} s kernel {sol.ms:.1f} ms  status {np.bincount(sol.status, minlength=5)} "
          f"iters {sol.iterations.min()}..{sol.iterations.max()} kkt max {sol.kkt_error.max():.2e}  NLP/s {B / (sol.ms * 1e-3):.1f}", flush=True)
    return s_tv, sol


s_tv, sol = run(int(os.environ.get("B1", 64)), int(os.environ.get("N1", 100)), "short")
emu = os.path.join(ROOT, "tests", "host_emu", "_build", "libk7emu.so")
if os.path.exists(emu):
    L = C.CDLL(emu)
    dp = C.POINTER(C.c_double)
    t = nlp.nlp_tables(r, p)
    o = nlp.NlpOpts(); nlp.api.load_library().eepacc_nlp_default_opts(C.byref(o))
    worst = 0.0
    for b in range(3):
        z = np.zeros(sol.z.shape[1]); st = np.zeros(9)
        row = np.ascontiguousarray(s_tv[b])
        L.k7_emu_solve(C.byref(p), C.byref(r), C.byref(t), row.ctypes.data_as(dp), sol.z.shape[1] // 10, C.byref(o), None,
                       z.ctypes.data_as(dp), st.ctypes.data_as(dp))
        worst = max(worst, float(np.abs(z - sol.z[b]).max()))
        print(f"  scenario {b}: gpu f {sol.f[b]:.6e} emu f {st[8]:.6e} iters gpu {sol.iterations[b]} emu {int(st[1])} max|z gpu - z emu| {np.abs(z - sol.z[b]).max():.3e}", flush=True)
print(f"elapsed {time.time() - t00:.1f} s", flush=True)
if os.environ.get("FULL", "1") == "1":
    g = np.load(os.path.join(ROOT, "tests", "golden", "nlp_mc.npz"))
    s_tv, sol = run(int(os.environ.get("B2", 32)), 870, "full cycle")
    print(f"  scenario 0 vs savedNLPsol.mat: max|s| {np.abs(sol.s_opt[0] - g['s_opt']).max():.3e} m  max|v| {np.abs(sol.v_opt[0] - g['v_opt']).max():.3e} m/s  "
          f"max|Fm| {np.abs(sol.Fm_opt[0] - g['Fm_opt']).max():.3e} N  f {sol.f[0]:.10e}  (ipopt tSolve {float(g['tSolve']):.0f} s)", flush=True)
print(f"elapsed {time.time() - t00:.1f} s", flush=True)
