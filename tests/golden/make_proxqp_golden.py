"""Generates tests/golden/proxqp_{mc,abo}_ab.npz -- an INDEPENDENT pin of the QP oracle.

Run in the development container only (needs /root/reference for the vendored proxsuite headers):
    make -C oracle ref && python tests/golden/make_proxqp_golden.py

For closed-loop steps 0, 8, .., 40 of the golden ABMPC run (both trees) the dense QP that the
oracle built (its H and G are pinned to the reference's stored H/G at 1e-12, test_oracle_golden.py) is
re-solved COLD by ProxQP 0.3.2 (casadi_3_6_3/include/proxsuite, a different algorithm) at eps_abs=1e-11.
Stored per step: ProxQP's objective, its minimiser x (first stage + full vector norm), residuals, and the
set of rows whose multiplier is non-zero (|z| > 1e-7: the strictly active set of the unique optimum).
The CPU test tests/test_oracle_proxqp.py then holds the oracle to these values without the reference.
FBMPC is not covered: its Hessian is indefinite and ProxQP does not converge on it.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from eepacc_mpc_casadi_matlab_b200 import settings as S, usecase as U   # noqa: E402
from oracle import oracle as O                                           # noqa: E402

dp = C.POINTER(C.c_double)
P = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libproxqp_ref.so"))
P.proxqp_ref_solve.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, dp, C.c_double, C.c_int, dp, dp, dp, dp, dp,
                               C.POINTER(C.c_int)]


def proxqp(H, g, A, lba, uba, eps=1e-11):
    n, m = len(g), len(lba)
    Hc, Ac = np.asfortranarray(H), np.asfortranarray(A)
    x, z = np.zeros(n), np.zeros(m)
    cost, pr, du = np.zeros(1), np.zeros(1), np.zeros(1)
    it = C.c_int(0)
    st = P.proxqp_ref_solve(n, m, Hc.ctypes.data_as(dp), g.ctypes.data_as(dp), Ac.ctypes.data_as(dp),
                            lba.ctypes.data_as(dp), uba.ctypes.data_as(dp), eps, 20000,
                            x.ctypes.data_as(dp), z.ctypes.data_as(dp), cost.ctypes.data_as(dp),
                            pr.ctypes.data_as(dp), du.ctypes.data_as(dp), C.byref(it))
    return st, x, z, cost[0], pr[0], du[0], it.value


STRIDE = 8        # ProxQP at eps 1e-11 needs minutes per QP (10-17 k outer iterations): a sample of the steps
LAST = 40         # ... 0, 8, .., 40 (acceleration from standstill behind the lead: the working set changes every step)


def main():
    d = np.load(os.path.join(ROOT, "tests", "golden", "lead_TO01_EAD.npz"))
    lead = np.stack([d["s_tv"], d["v_tv"]], 1)
    for tree in ("mc", "abo"):
        p = S.default_params(tree)
        V = O.Vehicle(S.CTRL_NAMES["ab"], p, U.default_route())
        n_steps = 871
        cost, x1, xn, res, act, stat = [], [], [], [], [], []
        worst_c, worst_x, ws_mismatch = 0.0, 0.0, 0
        for k in range(n_steps):
            out = V.step(*lead[k])
            if k > LAST:
                break
            if k % STRIDE:
                continue
            H, g, A, lba, uba = O.qp_arrays(V.qp())
            st, x, z, c, pr, du, it = proxqp(H, g, A, lba, uba)
            stat.append(st); cost.append(c); x1.append(x[:5].copy()); xn.append(np.linalg.norm(x))
            res.append([pr, du, it])
            a = np.zeros(len(lba), dtype=np.int8)
            a[z > 1e-7] = 1; a[z < -1e-7] = -1          # upper / lower bound strictly active
            act.append(np.packbits(a != 0))
            worst_c = max(worst_c, abs(out[11] - c) / (1 + abs(c)))
            ws = V.working_set()[:len(lba)]               # oracle working set over the m rows (far bounds follow)
            # strictly active rows must be in the oracle's working set
            ws_mismatch += int(np.sum((a != 0) & (ws == 0)))
            if True:
                print(tree, k, "status", st, "cost", c, "oracle", out[11], "iters", it, "res", pr, du, flush=True)
        print(f"{tree}: worst rel objective gap oracle vs ProxQP = {worst_c:.3e}; "
              f"strictly-active rows missing from the oracle working set: {ws_mismatch}; "
              f"ProxQP failures: {int(np.sum(np.array(stat) != 0))}")
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"proxqp_{tree}_ab.npz"),
                            cost=np.array(cost), x_first=np.array(x1), x_norm=np.array(xn),
                            residuals=np.array(res), active_bits=np.array(act), status=np.array(stat),
                            steps=np.arange(0, LAST + 1, STRIDE),
                            meta=np.array(["proxsuite 0.3.2 dense, cold start, eps_abs=1e-11, eps_rel=0"]))


if __name__ == "__main__":
    main()
