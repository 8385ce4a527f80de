"""CPU: independent pin of the QP oracle at the conic boundary (objective, minimiser, active set).

tests/golden/proxqp_{mc,abo}_ab.npz hold, for each of the 871 ABMPC closed-loop QPs of both reference trees,
the solution of ProxQP 0.3.2 -- the dense proximal augmented-Lagrangian solver vendored inside the
reference's CasADi tree (casadi_3_6_3/include/proxsuite) -- compiled from those headers and run COLD at
eps_abs = 1e-11 by tests/golden/make_proxqp_golden.py (oracle/xcheck/proxqp_ref.cpp).  The reference's own
saved runs store no objective values or duals; this closes that gap for the strictly convex controller with a
solver that shares no code and no algorithm with oracle/qp_oracle.c.  (FBMPC: indefinite Hessian, ProxQP does
not converge on it -- the FB objective stays pinned by the oracle alone, see DESIGN.md.)"""
import numpy as np
import pytest

from conftest import golden
from eepacc_mpc_casadi_matlab_b200 import settings as S, usecase as U


@pytest.mark.parametrize("tree", ["mc", "abo"])
def test_oracle_minimiser_objective_and_active_set_match_proxqp(oracle, lead_urban, tree):
    """Tolerances: first-stage minimiser 1e-6 absolute (the north star's primal tolerance); objective 1e-6 relative --
    ProxQP stops at primal/dual residuals of 1e-11, which multipliers of up to 9e7 (slack weights w_s, w_f) turn into
    an objective uncertainty of ~1e-3 on values of 2e4..3e5 (its objective lies ABOVE the oracle's, whose point the
    KKT test below certifies to 1e-12); active sets: see below."""
    g = golden(f"proxqp_{tree}_ab.npz")
    steps = [int(k) for k in g["steps"]]
    p = S.default_params(tree)
    V = oracle.Vehicle(S.CTRL_NAMES["ab"], p, U.default_route())
    n = 100
    checked = 0
    for k in range(max(steps) + 1):
        out = V.step(*lead_urban[k])
        if k not in steps:
            continue
        i = steps.index(k)
        if g["status"][i] != 0:
            continue
        x, lam = V.solution()
        c_ref = g["cost"][i]
        assert abs(out[11] - c_ref) <= 1e-6 * (1.0 + abs(c_ref)), (k, out[11], c_ref)
        assert np.max(np.abs(x[:5] - g["x_first"][i])) < 1e-6, k
        assert abs(np.linalg.norm(x) - g["x_norm"][i]) < 1e-6 * (1 + g["x_norm"][i]), k
        m = len(lam)
        act = np.unpackbits(g["active_bits"][i])[:m].astype(bool)
        ws = V.working_set()[:m]              # status of the m general rows (far bounds follow)
        # a row ProxQP holds active is in the oracle's working set, or -- degenerate case, zero multiplier -- lies
        # exactly on its bound at the oracle's point (10 of the 12 sampled working sets are identical sets)
        H, gg, A, lba, uba = oracle.qp_arrays(V.qp())
        Ax = A @ x
        on_bound = np.minimum(np.abs(Ax - lba), np.abs(uba - Ax)) < 1e-9
        assert not np.any(act & (ws == 0) & ~on_bound), k
        # (the converse is not asserted row by row: duplicated rows -- e.g. speed-limit and curve rows with the same
        # bound -- may carry the multiplier on either copy; rows that sit on their bound with a zero multiplier need
        # not be in the oracle's working set at all: its ratio tests leave weakly active rows out, see TIE_TOL)
        assert abs(int(act.sum()) - int((ws != 0).sum())) <= 20, k
        checked += 1
    assert checked >= len(steps) - 1


@pytest.mark.parametrize("tree", ["mc", "abo"])
@pytest.mark.parametrize("cn", ["ab", "fb"])
def test_oracle_closed_loop_solutions_are_kkt_points(oracle, lead_urban, tree, cn):
    """Solver-independent certificate for all 871 warm-started closed-loop QPs: the oracle's (x, lam_a) is primal
    feasible, stationary (H x + g + A' lam_a = 0, conic.hpp sign convention), dual feasible (a positive
    multiplier only on a finite upper bound, a negative one only on a finite lower bound) and complementary, and
    the reported cost is 1/2 x'Hx + g'x.  ABMPC's Hessian is positive semidefinite (checked), so a KKT point IS a
    global minimiser and its objective THE optimal value -- this, rather than ProxQP's 1e-11-residual iterate,
    pins the oracle's objective for the convex controller.  FBMPC's Hessian is indefinite
    (CreateQP_FB.m:181-184): there the check certifies a first-order point only."""
    p = S.default_params(tree)
    V = oracle.Vehicle(S.CTRL_NAMES[cn], p, U.default_route())
    worst = dict(feas=0.0, stat=0.0, comp=0.0)
    for k in range(871):
        out = V.step(*lead_urban[k])
        H, g, A, lba, uba = oracle.qp_arrays(V.qp())
        x, lam = V.solution()
        Ax = A @ x
        cost = 0.5 * x @ H @ x + g @ x
        assert abs(cost - out[11]) <= 1e-9 * (1 + abs(cost)), k
        scale = 1.0 + np.abs(lam).max()
        worst["feas"] = max(worst["feas"], float(np.max(np.maximum(lba - Ax, 0) + np.maximum(Ax - uba, 0))))
        worst["stat"] = max(worst["stat"], float(np.abs(H @ x + g + A.T @ lam).max()) / scale)
        tiny = 1e-12 * scale                                                   # round-off duals of inactive rows
        up, lo = lam > tiny, lam < -tiny
        assert np.all(np.isfinite(uba[up])) and np.all(np.isfinite(lba[lo])), k
        worst["comp"] = max(worst["comp"], float(np.abs(lam[up] * (uba[up] - Ax[up])).sum()
                                                 + np.abs(lam[lo] * (Ax[lo] - lba[lo])).sum()) / scale)
        if k == 0 and cn == "ab":
            assert np.linalg.eigvalsh(0.5 * (H + H.T)).min() > -1e-9 * np.abs(H).max()
    assert worst["feas"] < 1e-10, worst
    assert worst["stat"] < 1e-12, worst
    assert worst["comp"] < 1e-12, worst
