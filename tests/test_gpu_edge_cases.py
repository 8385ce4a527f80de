"""GPU: edge cases of the C-ABI and of the scenario inputs."""
import ctypes as C

import numpy as np
import pytest

from eepacc_mpc_casadi_matlab_b200 import settings as S, usecase as U

pytestmark = pytest.mark.gpu
CTRL = {"ab": S.AB, "fb": S.FB, "bl": S.BL, "tv": S.TV}


def _api():
    from eepacc_mpc_casadi_matlab_b200 import api
    return api


@pytest.mark.parametrize("cn", ["ab", "fb"])
def test_route_with_stops_lights_curves_and_slopes(oracle, lead_urban, cn):
    """Maximum-size style route blocks (stops, traffic lights, curves, hills): every row family of
    EstimateRouteAndComfortBounds is exercised; CUDA follows the oracle step by step."""
    api = _api()
    p = S.default_params("mc")
    r = U.generate_use_case([[50, 0], [30, 250], [70, 600], [50, 1500]],
                            curves=[[-1 / 12, 120, 140], [1 / 30, 400, 460], [-1 / 9, 900, 905]],
                            slopes=[[4, 100, 300], [-3, 500, 700]], stops=[200.0, 820.0],
                            tls=[[350.0, 2.0, 12.0, 10.0], [1100.0, 7.0, 15.0, 12.0]])
    lead = lead_urban[:160].copy()
    lead[:, 0] += 60.0                       # give the ego room so that the route rows bind
    eng = api.Engine(cn, p, r, B=1)
    V = oracle.Vehicle(CTRL[cn], p, r)
    seen_stop = False
    for k in range(160):
        out = eng.step(lead[k:k + 1])
        o = V.step(*lead[k])
        assert out["status"][0] == o[9], k
        assert np.max(np.abs(out["z1"][0, :2] - o[:2])) < 1e-6, k
        assert np.max(np.abs(out["z1"][0, 4:] - o[4:8])) < 1e-5, k
        seen_stop |= o[0] > 100.0
    assert seen_stop
    eng.close()


def test_batch_of_one_and_odd_batch_sizes(oracle, lead_urban):
    api = _api()
    p = S.default_params("mc")
    y = np.r_[0.0, np.diff(lead_urban[:, 0]) / p.Ts][:871]
    for B in (1, 3, 149):
        routes, lead = U.perturbed_scenarios(B, y, n_steps=40)
        eng = api.Engine("ab", p, routes, B=B)
        eng.set_lead(lead)
        log, kpi = eng.run(40)
        eng.close()
        assert log.shape == (40, B, 12) and kpi.shape == (B, S.N_KPI)
        b = B - 1
        ref, kref = oracle.run(S.AB, p, routes[b], lead[:, b, :])
        assert np.max(np.abs(log[:, b, :2] - ref[:, :2])) < 1e-6
        assert np.allclose(kpi[b], kref, rtol=1e-9, atol=1e-9)


def test_argument_errors_are_reported_not_crashes():
    api = _api()
    p = S.default_params("mc")
    eng = api.Engine("ab", p, U.default_route(), B=2)
    with pytest.raises(api.EepaccError):                       # batch must match the mem's batch
        eng._call("eepacc_step_ab", [np.zeros((3, 2))], [np.zeros((3, 8)), np.zeros(3), np.zeros(3), np.zeros(3)], 3)
    with pytest.raises(api.EepaccError):                       # FB entry point on an AB mem
        eng._call("eepacc_step_fb", [np.zeros((2, 2))], [np.zeros((2, 8)), np.zeros(2), np.zeros(2), np.zeros(2)], 2)
    with pytest.raises(api.EepaccError):                       # run without a lead trace
        eng.run(5)
    eng.close()
    bad = S.default_params("mc")
    bad.N = 10
    with pytest.raises(api.EepaccError):
        api.Engine("ab", bad, U.default_route(), B=1)


def test_qpsolve_reports_infeasible_and_handles_equalities(oracle):
    """Conic-scheme semantics at the edges: inconsistent constraints -> status 2 (x still returned,
    error_on_fail=false); equality rows (lba == uba) are kept in the working set."""
    api = _api()
    p = S.default_params("mc")
    eng = api.Engine("tv", p, U.default_route(), B=2)           # n = 40, m = 220
    n, m = eng.n, eng.m
    rng = np.random.default_rng(3)
    H = np.stack([np.eye(n) * 2.0] * 2)
    g = rng.normal(size=(2, n))
    A = np.zeros((2, m, n)); lba = np.full((2, m), -np.inf); uba = np.full((2, m), np.inf)
    # vehicle 0: x0 + x1 = 1 (equality), x2 >= 0.5 ; vehicle 1: x0 >= 1 and x0 <= 0 (infeasible)
    A[0, 0, 0] = A[0, 0, 1] = 1; lba[0, 0] = uba[0, 0] = 1.0
    A[0, 1, 2] = 1; lba[0, 1] = 0.5
    A[1, 0, 0] = 1; lba[1, 0] = 1.0
    A[1, 1, 0] = 1; uba[1, 1] = 0.0
    res = eng.qpsolve(H, g, A, lba, uba)
    assert res["status"][0] == 0 and res["status"][1] == 2
    x = res["x"][0]
    assert abs(x[0] + x[1] - 1.0) < 1e-10 and x[2] >= 0.5 - 1e-10
    qs = oracle.QPSolver(n, m)
    ro = qs.solve(H[0], g[0], A[0], lba[0], uba[0])
    assert np.max(np.abs(x - ro["x"])) < 1e-9 and abs(res["cost"][0] - ro["cost"]) < 1e-9
    eng.close()


def test_device_pointer_mode_matches_host_mode(lead_urban):
    torch = pytest.importorskip("torch")
    api = _api()
    p = S.default_params("mc")
    r = U.default_route()
    a, b = api.Engine("fb", p, r, B=1), api.Engine("fb", p, r, B=1)
    for k in range(12):
        h = a.step(lead_urban[k:k + 1])
        d = b.step(torch.from_numpy(lead_urban[k:k + 1].copy()).cuda())
        assert np.array_equal(h["z1"], d["z1"].cpu().numpy())
    a.close(); b.close()


def test_two_devices_in_one_process(lead_urban):
    """The dynamic shared-memory opt-in of the kernels is per device: an Engine on device 1 created after one on
    device 0 has run must work (it used to be cached process-wide), give the same bits, and leave the caller's
    current device alone."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    api = _api()
    p = S.default_params("mc")
    y = np.r_[0.0, np.diff(lead_urban[:, 0]) / p.Ts][:871]
    routes, lead = U.perturbed_scenarios(40, y, n_steps=30)
    logs = []
    torch.cuda.set_device(0)
    for dev in (0, 1, 0):
        eng = api.Engine("fb", p, routes, B=40, device=dev)
        assert eng.device_index == dev
        eng.set_lead(lead)
        log, _ = eng.run(30)
        logs.append(log.copy())
        eng.close()
        assert torch.cuda.current_device() == 0
    assert np.array_equal(logs[0], logs[1]) and np.array_equal(logs[0], logs[2])
