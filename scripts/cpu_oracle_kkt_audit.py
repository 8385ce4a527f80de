"""Developer aid (CPU only): KKT audit of the ORACLE's closed-loop solutions on the bench batch: for every QP of a
scenario, are the multipliers of the final working set sign-feasible (Hx + g + A'lam = 0, lam <= 0 on rows active at
their lower bound, >= 0 at their upper bound)?  A wrong sign means the active-set path ended on a non-optimal vertex.

    python scripts/cpu_oracle_kkt_audit.py [ctrl] [n_scen] [n_steps] [first]
"""
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(args):
    ctrl, b, ns = args
    import bench
    from eepacc_mpc_casadi_matlab_b200 import settings as S
    from oracle import oracle as O
    p, routes, lead = bench.make_workload(ctrl, 1, b, ns)
    V = O.Vehicle(S.CTRL_NAMES[ctrl], p, routes[0])
    bad = []
    for k in range(ns):
        out = V.step(*lead[k, 0])
        H, g, A, lba, uba = O.qp_arrays(V.qp())
        x, lam = V.solution()
        Ax = A @ x
        tol = 1e-8 * (1 + np.abs(Ax))
        al = (np.abs(Ax - lba) < tol) & ~(np.abs(Ax - uba) < tol)
        au = (np.abs(Ax - uba) < tol) & ~(np.abs(Ax - lba) < tol)
        sc = 1e-7 * (1 + np.abs(lam).max())
        wrong = max(lam[al].max(initial=0.0), -lam[au].min(initial=0.0))
        stat = np.abs(H @ x + g + A.T @ lam).max() / (1 + np.abs(g).max())
        if wrong > sc or stat > 1e-9 or out[9] != 0:
            bad.append((k, float(wrong), float(stat), int(out[9])))
    return b, bad


if __name__ == "__main__":
    ctrl = sys.argv[1] if len(sys.argv) > 1 else "fb"
    nsc = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    ns = int(sys.argv[3]) if len(sys.argv) > 3 else 871
    first = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    from oracle import oracle as O
    O.build()
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        res = pool.map(worker, [(ctrl, first + b, ns) for b in range(nsc)], chunksize=1)
    tot = 0
    for b, bad in res:
        tot += len(bad)
        if bad:
            print(f"scen {b}: {len(bad)} QPs end with a wrong-sign multiplier / residual: first {bad[:4]}")
    print(f"{tot} of {nsc * ns} oracle solutions are not KKT points ({ctrl}, scenarios {first}..{first + nsc - 1})")
