"""Developer aid (CPU only): run two arithmetic variants of the oracle (strict IEEE vs FMA-contracted -O3)
over the bench batch and report the first step at which their closed loops part.  A step where they part is
a step whose active-set path is decided by round-off -- the same steps the CUDA solver (different summation
order again) cannot be expected to reproduce.  Used to find and remove noise-decided branches.

    python scripts/cpu_oracle_sensitivity.py [ctrl] [n_scen] [n_steps] [first]
"""
import ctypes as C
import multiprocessing as mp
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ALT = "/tmp/orc/liboracle_fma.so"


def build_alt():
    os.makedirs(os.path.dirname(ALT), exist_ok=True)
    src = [os.path.join(ROOT, "oracle", f) for f in ("eepacc_oracle.c", "qp_oracle.c")]
    subprocess.check_call(["gcc", "-O3", "-march=native", "-ffp-contract=fast", "-fPIC", "-std=c11", "-shared",
                           "-o", ALT] + src + ["-lm"])


def run_lib(path, ctrl_id, p, route, lead):
    from eepacc_mpc_casadi_matlab_b200.settings import N_KPI, Params, Route
    L = C.CDLL(path)
    dp = C.POINTER(C.c_double)
    L.orc_run.argtypes = [C.c_int, C.POINTER(Params), C.POINTER(Route), dp, C.c_int, C.c_int, dp, dp]
    lead = np.ascontiguousarray(lead)
    n = lead.shape[0]
    log = np.zeros((n, 12)); kpi = np.zeros(N_KPI)
    L.orc_run(ctrl_id, C.byref(p), C.byref(route), lead.ctypes.data_as(dp), 2, n, log.ctypes.data_as(dp),
              kpi.ctypes.data_as(dp))
    return log, kpi


def worker(args):
    ctrl, b, ns = args
    import bench
    from eepacc_mpc_casadi_matlab_b200 import settings as S
    from oracle import oracle as O
    p, routes, lead = bench.make_workload(ctrl, 1, b, ns)
    a, ka = run_lib(O.build(), S.CTRL_NAMES[ctrl], p, routes[0], lead[:, 0, :])
    f, kf = run_lib(ALT, S.CTRL_NAMES[ctrl], p, routes[0], lead[:, 0, :])
    ds = np.abs(a[:, :2] - f[:, :2]).max(axis=1)
    dobj = np.abs(a[:, 11] - f[:, 11]) / (1 + np.abs(a[:, 11]))
    bad = np.nonzero((ds > 1e-7) | (dobj > 1e-8) | (a[:, 9] != f[:, 9]) | (a[:, 10] != f[:, 10]))[0]
    first = int(bad[0]) if len(bad) else -1
    it_diff = int((a[:, 10] != f[:, 10]).sum())
    return b, first, float(ds.max()), float(dobj.max()), it_diff, float(a[:, 1].min()), int((a[:, 9] != 0).sum())


if __name__ == "__main__":
    ctrl = sys.argv[1] if len(sys.argv) > 1 else "fb"
    nsc = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    ns = int(sys.argv[3]) if len(sys.argv) > 3 else 871
    first = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    build_alt()
    from oracle import oracle as O
    O.build()
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        res = pool.map(worker, [(ctrl, first + b, ns) for b in range(nsc)], chunksize=1)
    nb = 0
    for b, fs, ds, dobj, itd, vmin, nfail in res:
        flag = "" if fs < 0 else "  <-- DIVERGES"
        nb += fs >= 0
        print(f"scen {b:5d}: first diff step {fs:4d}  max|ds| {ds:.2e}  max rel dobj {dobj:.2e}  iter diffs {itd:3d}  failed {nfail}{flag}")
    print(f"{nb} of {nsc} scenarios diverge between the two arithmetic variants")
