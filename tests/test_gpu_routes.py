"""GPU: route previews built on the device (k5_route_gen_kernel) against the host statement of
GenerateUseCase.m:50-107 + SimplifyPWA.m:14-49 (usecase.generate_use_case) -- bit for bit, quirks included."""
import ctypes as C

import numpy as np
import pytest

from eepacc_mpc_casadi_matlab_b200 import settings as S, usecase as U

pytestmark = pytest.mark.gpu


def _api():
    from eepacc_mpc_casadi_matlab_b200 import api
    return api


def _bytes(route):
    return bytes(C.string_at(C.addressof(route), C.sizeof(S.Route)))


def _same(dev, host):
    """Route blocks are equal in everything the kernels read: counts, the first n points of every PWA, stops, lights, mass."""
    for n_name, fields in (("n_sl", ("s_speedLim", "v_speedLim")), ("n_curv", ("s_curv", "curvature")),
                           ("n_slope", ("s_slope", "slope")), ("n_stop", ("stopLoc",))):
        n = getattr(host, n_name)
        if getattr(dev, n_name) != n:
            return False, (n_name, getattr(dev, n_name), n)
        for f in fields:
            a, b = list(getattr(dev, f))[:n], list(getattr(host, f))[:n]
            if a != b:
                return False, (f, a, b)
    if dev.n_tl != host.n_tl or [list(r) for r in dev.TLLoc][:host.n_tl] != [list(r) for r in host.TLLoc][:host.n_tl]:
        return False, ("TLLoc",)
    return dev.mass_scale == host.mass_scale, ("mass_scale",)


def test_device_routes_equal_the_host_statement_on_the_named_use_cases():
    api = _api()
    p = S.default_params("mc")
    nums = sorted(U.USE_CASES)
    specs, hosts = [], []
    for num in nums:
        uc = U.USE_CASES[num]
        args = dict(speed_lim_zones=uc.get("speed"), curves=uc.get("curves"), slopes=uc.get("slopes"),
                    stops=uc.get("stops", ()), tls=uc.get("tls", ()))
        specs.append(U.route_spec(mass_scale=1.0 + 0.01 * num, **args))
        hosts.append(U.generate_use_case(mass_scale=1.0 + 0.01 * num, **args))
    eng = api.Engine("ab", p, None, B=len(nums))
    eng.gen_routes(specs)
    dev = eng.fetch_routes()
    eng.close()
    for i, num in enumerate(nums):
        ok, why = _same(dev[i], hosts[i])
        assert ok, (num, why)
        assert _bytes(dev[i]) == _bytes(hosts[i]), num           # and the unused tails are zero on both sides


def test_device_routes_equal_the_host_statement_on_the_sweep_table_and_drive_the_same_closed_loop():
    api = _api()
    p = S.default_params("mc")
    B = 384
    pert, specs = U.spec_table_fast(B, first=1000, full_route=True, mass=True)
    hosts = (S.Route * B)()
    for b in range(B):
        r = U.spec_to_route(specs[b])
        C.memmove(C.byref(hosts[b]), C.byref(r), C.sizeof(S.Route))
    eng = api.Engine("bl", p, None, B=B)
    eng.gen_routes(specs)
    dev = eng.fetch_routes()
    for b in range(B):
        ok, why = _same(dev[b], hosts[b])
        assert ok, (b, why)
    assert sum(dev[b].n_curv > 4 for b in range(B)) > B // 2 and sum(dev[b].n_stop > 0 for b in range(B)) > B // 2
    d = np.load(__import__("os").path.join(__import__("bench").GOLDEN, "lead_TO01_EAD.npz"))
    y = np.r_[0.0, np.diff(d["s_tv"]) / p.Ts][:871]
    lead = U.synth_lead_host(pert, y, None, 60)
    eng.set_lead(lead)
    log_d, kpi_d = eng.run(60)
    eng.close()
    eng = api.Engine("bl", p, hosts, B=B)
    eng.set_lead(lead)
    log_h, kpi_h = eng.run(60)
    eng.close()
    assert np.array_equal(log_d, log_h) and np.array_equal(kpi_d, kpi_h)


def test_route_that_does_not_fit_the_block_is_reported():
    api = _api()
    p = S.default_params("mc")
    curves = [[1 / (10.0 + k), 100 + 40 * k, 120 + 40 * k] for k in range(12)]       # 12 curves x 4 points > 48
    with pytest.raises(ValueError):
        U.generate_use_case([[60, 0]], curves)
    eng = api.Engine("ab", p, None, B=2)
    with pytest.raises(api.EepaccError, match="scenario 1 needs more curvature points"):
        eng.gen_routes([U.route_spec([[60, 0]]), U.route_spec([[60, 0]], curves)])
    eng.close()
