"""GPU: the reference's named use cases (GetUseCase.m 1-12) through the C-ABI closed loop, against the oracle.

Covers what use case 0 never reaches: non-zero initial speed, stops, traffic lights, curves, a hill, a cut-in,
recorded lead traces and runs without a target vehicle (s_tv = inf).  Criteria as in test_gpu_parity.py:
ABMPC and FBMPC must follow the oracle over the whole run with the same solver status and the same number of
working-set changes at every step (FBMPC also from a non-zero initial speed, where RunOpt_FBMPC.m:84 makes the first
QPs infeasible: the kernel must return what the oracle returns); for the BL/TV LPs a scenario is compared up to its
first divergent step, which must follow a step solved to the same objective."""
import numpy as np
import pytest

from eepacc_mpc_casadi_matlab_b200 import report as R, settings as S, usecase as U

pytestmark = pytest.mark.gpu
CTRL = {"ab": S.AB, "fb": S.FB, "bl": S.BL, "tv": S.TV}
MAX_STEPS = 300
FB_CASES = list(range(1, 13))
# Use case 11 (stop lines at 80/450/600 m, 300 steps): the closed loop itself amplifies a state difference by ~1e7
# between step 40 and step 300 (the stop funnel v_stop_max = |stop - s| slope + v_stop feeds the position error
# straight back; measured on the CPU between two builds of the ORACLE that differ only in FMA contraction: 1e-13 m
# at step 40, 1e-10 at 60, 1e-6 at 212, 4e-6 at 218).  No two arithmetics can stay within 1e-6 there; the test keeps
# the per-step decisions exact (status, working-set changes) and bounds the states by that measured growth.
FB_STATE_TOL = {11: 1e-4}


def _api():
    from eepacc_mpc_casadi_matlab_b200 import api
    return api


def _lead_for(oracle, num, p, uc):
    if uc["lead"] is not None:
        return uc["lead"]

    def run_tv(pp, route, n_steps):
        log, _ = oracle.run(S.TV, pp, route, np.zeros((n_steps, 2)))
        return log[:, 0], log[:, 1]
    return U.generate_lead_tvmpc(run_tv, p, uc["route"], uc["n_steps"])


def _compare(cn, log, ref, strict, w_max=0.0, tol_state=None, obj_floor=1e-10):
    """w_max: largest penalty weight of the controller.  The objective is dominated by weight * slack terms, so a
    1e-11 difference in a slack (states agree to 1e-6, the bar) moves it by 1e-11 * w_max: the 1e-9 relative bar
    on the objective is applied up to that absolute floor."""
    lp = cn in ("bl", "tv")
    tol = tol_state or (2e-6 if lp else 1e-6)
    loose = tol_state is not None
    prev_rel, compared = 0.0, 0
    for k in range(len(ref)):
        o, g = ref[k], log[k]
        ok = np.max(np.abs(g[:2] - o[:2])) < tol and g[9] == o[9]
        if not ok and not strict:
            assert prev_rel < 1e-3, (k, "diverged after a step whose objectives differ", prev_rel)
            break
        assert ok, (k, g[:2], o[:2], g[9], o[9])
        prev_rel = abs(g[11] - o[11]) / (1 + abs(o[11]))
        if strict:
            assert g[10] == o[10], (k, "working-set changes", g[10], o[10])
            if not loose:
                assert prev_rel <= 1e-9 or abs(g[11] - o[11]) <= obj_floor * w_max, (k, "objective", g[11], o[11])
                assert abs((g[2] + g[3]) - (o[2] + o[3])) < 1e-4, (k, "Fm+Fb")
                assert np.max(np.abs(g[4:8] - o[4:8])) < 1e-5, (k, "slacks")
        compared += 1
    return compared


@pytest.mark.parametrize("num", list(range(1, 13)))
def test_abmpc_follows_oracle_on_every_use_case(oracle, num):
    api = _api()
    p = S.default_params("mc")
    uc = U.get_use_case(num, p)
    n = min(uc["n_steps"], MAX_STEPS)
    lead = _lead_for(oracle, num, p, uc)[:n]
    sol = api.RunOpt_ABMPC(p, uc["route"], lead)
    ref, kref = oracle.run(S.AB, p, uc["route"], lead)
    log = np.stack([sol[k][:, 0] for k in ("s_opt", "v_opt", "Fm_opt", "Fb_opt", "xi_v_opt", "xi_h_opt", "xi_s_opt",
                                            "xi_f_opt", "a_opt")] +
                   [sol["exitMessage"][:, 0].astype(float), sol["iter"][:, 0], sol["cost"][:, 0]], axis=1)
    ref[:, 9] = ref[:, 9] != 0                                       # exitMessage is a flag: compare failure yes/no
    assert _compare("ab", log, ref, strict=True, w_max=max(p.W_AB)) == n
    assert np.allclose(sol["kpi"][0], kref, rtol=1e-6, atol=1e-6)
    rep = R.kpi_report(sol, p, uc["route"], lead=lead[:, None, :])    # host report == device KPI block
    assert rep["energy_kWh"][0] * 3.6e6 == pytest.approx(sol["kpi"][0, 0], rel=1e-9)
    assert rep["travel_time_at_cutoff_s"][0] == sol["kpi"][0, 2]
    assert rep["a_rms"][0] == pytest.approx(sol["kpi"][0, 6], rel=1e-9)


@pytest.mark.parametrize("cn", ["bl", "fb"])
def test_bl_and_fb_on_use_cases(oracle, cn):
    api = _api()
    total = compared = 0
    per_case = {}
    for num in (FB_CASES if cn == "fb" else range(1, 13)):
        p = S.default_params("mc")
        uc = U.get_use_case(num, p)
        n = min(uc["n_steps"], MAX_STEPS)
        lead = _lead_for(oracle, num, p, uc)[:n]
        eng = api.Engine(cn, p, uc["route"], B=1)
        eng.set_lead(lead[:, None, :])
        log, _ = eng.run(n)
        eng.close()
        ref, _ = oracle.run(CTRL[cn], p, uc["route"], lead)
        if cn == "fb":
            # objective floor: the FB objective is ~1e8 = w_s x slack; in a CLOSED loop the two runs' states drift apart
            # by ~1e-9 m (bar: 1e-6), which moves the slack by as much and the objective by 1e-8 * w_max (measured 0.5)
            c = _compare(cn, log[:, 0, :], ref, strict=True, w_max=max(p.W_FB), tol_state=FB_STATE_TOL.get(num), obj_floor=1e-7)
            assert c == n, (num, c, n)
        else:
            c = _compare(cn, log[:, 0, :], ref, strict=False)
        assert c >= 20, (num, c)
        per_case[num] = (c, n)
        total += n
        compared += c
    print(cn, "steps compared before the first alternative optimum, per use case:", per_case)
    assert compared >= (1.0 if cn == "fb" else 0.5) * total, (compared, total, per_case)


@pytest.mark.parametrize("num", [3, 4, 5, 6, 7, 11, 12])
def test_tvmpc_generates_the_target_vehicle(oracle, num):
    """Main.m:84-90 (commented RunOpt_TVMPC line): the TVMPC run on the GPU gives the lead trace the oracle gives."""
    api = _api()
    p = S.default_params("mc")
    uc = U.get_use_case(num, p)
    n = min(uc["n_steps"], MAX_STEPS)

    def run_gpu(pp, route, n_steps):
        sol = api.RunOpt_TVMPC(pp, route, n_steps)
        assert not sol["exitMessage"].any()
        return sol["s_opt"][:, 0], sol["v_opt"][:, 0]

    def run_cpu(pp, route, n_steps):
        log, _ = oracle.run(S.TV, pp, route, np.zeros((n_steps, 2)))
        return log[:, 0], log[:, 1]
    g = U.generate_lead_tvmpc(run_gpu, p, uc["route"], n)
    o = U.generate_lead_tvmpc(run_cpu, p, uc["route"], n)
    assert g[0].tolist() == [6.0, 0.0]
    bad = np.nonzero(np.max(np.abs(g - o), axis=1) >= 2e-6)[0]
    first = bad[0] if len(bad) else n                     # LP: alternative optima may part the two runs
    assert first >= 0.6 * n, (num, first, n)


@pytest.mark.parametrize("highway", [False, True])
def test_device_lead_synthesis_matches_host_statement(lead_urban, highway):
    """k5_lead_synth_kernel vs usecase.synth_lead_host on the same perturbation table: sin() is the only
    operation that is not correctly rounded on both sides, so traces agree to round-off (1e-9 m over 871 steps),
    the unperturbed scenario 0 bit for bit; the closed loop run on the synthesised traces equals the run on the
    uploaded ones."""
    from conftest import golden
    api = _api()
    p = S.default_params("mc")
    B, n = 300, 871
    y = np.r_[0.0, np.diff(lead_urban[:, 0]) / p.Ts][:871]
    w = golden("lead_WLTC.npz")
    hw = np.r_[0.0, np.diff(w["s_tv"]) / 0.5] if highway else None
    pert, routes = U.perturbation_table(B, highway=highway, mass=highway)
    host = U.synth_lead_host(pert, y, hw, n)
    eng = api.Engine("ab", p, routes, B=B)
    eng.synth_lead(pert, y, hw, n)
    dev = eng.fetch_lead()
    assert dev.shape == host.shape
    assert np.array_equal(dev[:, 0, :], host[:, 0, :])                     # scenario 0: no sin() involved
    assert np.max(np.abs(dev[:, :, 1] - host[:, :, 1])) < 1e-12
    assert np.max(np.abs(dev[:, :, 0] - host[:, :, 0])) < 1e-9
    assert (dev[:, :, 1] >= 0).all() and ((dev[:, :, 1] == 0) | (dev[:, :, 1] >= 0.1)).all()
    log_d, kpi_d = eng.run(60)
    eng.set_lead(host)
    log_h, kpi_h = eng.run(60)
    eng.close()
    assert np.max(np.abs(log_d[:, :, :2] - log_h[:, :, :2])) < 1e-6
    assert np.array_equal(log_d[:, :, 9], log_h[:, :, 9])


def test_device_lead_synthesis_edge_cases(lead_urban):
    """Shortest trace (2 steps), one vehicle, a ragged last CTA (130 vehicles), a delay longer than the run,
    and a base cycle shorter than the run (padded with its last sample, as the host statement does)."""
    api = _api()
    p = S.default_params("mc")
    y = np.r_[0.0, np.diff(lead_urban[:, 0]) / p.Ts][:871]
    for B, n, base in ((1, 2, y), (130, 15, y), (130, 120, y[:50]), (3, 871, y)):
        pert, routes = U.perturbation_table(B)
        if B > 2:
            pert[1, 1] = 20.0                     # 20 steps of standstill first (longer than the 15-step run)
            pert[2, 0] = 1.15
        host = U.synth_lead_host(pert, base, None, n)
        eng = api.Engine("bl", p, routes, B=B)
        eng.synth_lead(pert, base, None, n)
        dev = eng.fetch_lead()
        eng.close()
        assert dev.shape == (n, B, 2)
        assert np.max(np.abs(dev - host)) < 1e-9, (B, n)
        if B > 2 and n == 15:
            assert (dev[:, 1, 1] == 0).all() and (dev[:, 1, 0] == dev[0, 1, 0]).all()
    with pytest.raises(api.EepaccError):
        eng = api.Engine("bl", p, U.default_route(), B=1)
        try:
            eng.synth_lead(np.zeros((1, S.N_PERT)), y, None, 1)       # a trace needs at least two steps
        finally:
            eng.close()
