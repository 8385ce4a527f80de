"""GPU: the batched NLP benchmark (K7, csrc/k7_nlp.cu) through the C-ABI of include/eepacc_nlp.h, checked against
the oracle's restatement of RunOpt_NLP.m (feasibility, objective), the host emulation of the same device code and ipopt's
saved optimum.  Runs after the default-path tests (file name)."""
import ctypes as C

import numpy as np
import pytest

from conftest import golden
from eepacc_mpc_casadi_matlab_b200 import api, nlp, settings as S, usecase as U
from test_nlp_host_emu import emu, emu_solve, tables_of            # noqa: F401  (emu is a fixture)

pytestmark = pytest.mark.gpu


def scaled_leads(lead, B, N):
    sc = 1.0 + 0.1 * (np.arange(B) % 7) / 7.0
    return np.stack([6.0 + (lead[:N + 1] - 6.0) * s for s in sc])


def test_nlp_batch_matches_host_emulation_and_oracle(oracle, emu):              # noqa: F811
    """64 scenarios x 100 intervals (r2z4 probe: 0.67 s of kernel time on a B200, GPU == emulation to 3e-10)."""
    p, r = S.default_params("mc"), U.default_route()
    lead = golden("lead_TO01_EAD.npz")["s_tv"]
    B, N = 64, 100
    s_tv = scaled_leads(lead, B, N)
    sol = nlp.RunOpt_NLP(p, r, s_tv, n_steps=N)
    assert sol.ms > 0 and sol.z.shape == (B, 4 + 10 * N)
    assert set(sol.status) <= {0, 4}, np.bincount(sol.status)
    t = nlp.nlp_tables(r, p)
    sv, vv, ss, vs = tables_of(t)
    for b in range(0, B, 9):
        ev = oracle.nlp_eval(p, r, sv, vv, ss, vs, s_tv[b], sol.z[b])
        assert oracle.nlp_violation(ev, sol.z[b]) < 1e-9
        assert abs(ev["f"] - sol.f[b]) / ev["f"] < 1e-12
    for b in (0, 1, 5):
        z, st = emu_solve(emu, p, r, t, s_tv[b], N)
        assert abs(int(st[1]) - int(sol.iterations[b])) <= 5
        assert np.abs(z - sol.z[b]).max() < 1e-4
        assert abs(st[8] - sol.f[b]) / st[8] < 1e-10
    # scenarios 0 and 7 have the same lead trace (scale period 7): the batch position must not matter
    assert np.array_equal(sol.z[0], sol.z[7]) and sol.iterations[0] == sol.iterations[7]
    print(f"K7: {B} scenarios x {N} intervals in {sol.ms:.0f} ms = {B / sol.ms * 1e3:.0f} NLP/s, iterations {sol.iterations.min()}..{sol.iterations.max()}")


def test_nlp_full_cycle_reproduces_the_saved_ipopt_solution(oracle):
    """RunOpt_NLP.m's own case: 870 intervals of the urban cycle, one warp of scenarios (scenario 0 = the reference's).
    ipopt needed 620 s (savedNLPsol.mat:tSolve); bars as in tests/test_nlp_host_emu.py."""
    g = golden("nlp_mc.npz")
    p, r = S.default_params("mc"), U.default_route()
    lead = golden("lead_TO01_EAD.npz")["s_tv"]
    B, N = 32, 870
    s_tv = np.stack([lead[:N + 1]] * B)
    # odd scenarios: lead 0.75 % slower (converges on the host in 126 iterations; a 3 % slower lead puts a stage on a lookup
    # corner and runs into the iteration limit)
    s_tv[1:] = 6.0 + (s_tv[1:] - 6.0) * (1.0 - 0.0075 * (np.arange(1, B) % 2))[:, None]
    sol = nlp.RunOpt_NLP(p, r, s_tv, n_steps=N)
    assert sol.status[0] in (0, 4) and sol.iterations[0] < 200          # the reference's own case; the scaled ones are reported below
    t = nlp.nlp_tables(r, p)
    sv, vv, ss, vs = tables_of(t)
    ev = oracle.nlp_eval(p, r, sv, vv, ss, vs, lead, sol.z[0])
    assert oracle.nlp_violation(ev, sol.z[0]) < 1e-9
    assert abs(ev["f"] - sol.f[0]) / ev["f"] < 1e-12
    assert np.abs(sol.s_opt[0] - g["s_opt"]).max() < 1e-3
    assert np.abs(sol.v_opt[0] - g["v_opt"]).max() < 5e-5
    assert np.abs(sol.Fm_opt[0] - g["Fm_opt"]).max() < 1e-2
    assert np.abs(sol.j_opt[0] - g["j_opt"]).max() < 1e-5
    assert np.abs(sol.xi_v_opt[0] - g["xi_v_opt"]).max() < 5e-5
    print(f"K7 full cycle: {B} scenarios in {sol.ms / 1e3:.1f} s (ipopt: {float(g['tSolve']):.0f} s for one), status {np.bincount(sol.status, minlength=5)}, "
          f"max |s - ipopt| {np.abs(sol.s_opt[0] - g['s_opt']).max():.2e} m")


def test_nlp_argument_checks():
    p, r = S.default_params("mc"), U.default_route()
    sv = nlp.NlpSolver(p, r, 20, B=2)
    with pytest.raises(api.EepaccError):
        sv.solve(np.zeros((3, 21)))
    with pytest.raises(api.EepaccError):
        sv.solve(np.zeros((2, 21)), x0=np.zeros((2, 5)))
    with pytest.raises(api.EepaccError):
        sv.set_options(tol=-1.0)
    iw = (C.c_longlong * 2)(5, 0)
    assert sv.L.eepacc_nlp(None, None, iw, None, sv.mem) != 0
    sv.close()
