"""CPU: the LOGIC of the reduced-space QP kernel (csrc/k3r_qp.cu), compiled for the host as a single-thread
emulation (tests/host_emu/k3r_emu.cpp: NT = 1, barriers vanish), replayed on QP operands that K2 produced on a
B200 (tests/golden/qpdump_{ab,fb}_small.npz: structured block, g, bounds and the GPU kernel's answers for two
vehicles over the first 48 closed-loop steps, recorded by scripts/gpu_dump_qps.py).  Test infrastructure only:
the product never loads this library (there is no CPU fallback)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "host_emu"))
import replay as R  # noqa: E402


@pytest.mark.parametrize("cn,fc", [("ab", 48), ("ab", 24), ("fb", 48)])
def test_emulated_kernel_reproduces_recorded_gpu_answers(cn, fc):
    if R.cuda_include() is None:
        pytest.skip("cuda_runtime.h not found (host emulation needs the CUDA headers only)")
    d = np.load(os.path.join(HERE, "golden", f"qpdump_{cn}_small.npz"))
    n, m = int(d["n"]), int(d["m"])
    ns, B = d["x"].shape[:2]
    L = R.build()
    from eepacc_mpc_casadi_matlab_b200 import settings as S
    rp = R.Replayer(L, {"ab": S.AB, "fb": S.FB}[cn], n, m, B, fc)
    deferred = 0
    for k in range(ns):
        dl = rp.solve(d["sblk"][k], d["g"][k], d["lba"][k], d["uba"][k])
        deferred += len(dl)
        assert np.array_equal(rp.status, d["status"][k, :, 0]), k
        ec = np.abs(rp.cost - d["cost"][k, :, 0]) / (1 + np.abs(d["cost"][k, :, 0]))
        assert ec.max() < 1e-9, (k, ec.max())                       # objective: 1e-9 relative
        assert np.abs(rp.x - d["x"][k]).max() < 1e-6, k             # primal: 1e-6 absolute
    if fc < 40:
        assert deferred > 0          # small capacity: the hand-over path (second level / full space) is exercised


def test_estimator_modes_and_route_bounds_of_the_device_code_match_the_oracle(oracle):
    """EstimateVehicleTrajectory.m:55-88 in ALL modes (0 constant speed, 1 constant acceleration for tConstACC, 2 shifted
    previous solution) for the ego / lead / BL / TV settings, and EstimateRouteAndComfortBounds.m:60-199 on routes with
    speed zones, curves, stops, lights and slopes: the product's device functions (eepacc_dev.cuh, compiled for the host
    by tests/host_emu) against the oracle's literal restatement, bit for bit."""
    import ctypes as C
    from eepacc_mpc_casadi_matlab_b200 import settings as S, usecase as U
    L = R.build()
    if L is None:
        pytest.skip("no CUDA headers")
    O = oracle.lib()
    dp = C.POINTER(C.c_double)
    L.k3r_emu_est_traj.argtypes = [C.POINTER(S.Params), C.c_int] + [C.c_double] * 3 + [dp] * 4
    L.k3r_emu_bounds.argtypes = [C.POINTER(S.Params), C.POINTER(S.Route), C.c_int, dp, dp, C.c_double, dp]
    O.orc_est_traj.argtypes = [C.POINTER(S.Params), C.c_int] + [C.c_double] * 3 + [dp] * 4
    O.orc_bounds.argtypes = [C.POINTER(S.Params), C.POINTER(S.Route), C.c_int, dp, dp, C.c_double] + [dp] * 9
    rng = np.random.default_rng(3)
    N = 20
    checked = 0
    for mode in (0, 1, 2):
        for setting in (0, 1, 2, 3):
            p = S.default_params("mc")
            p.est_ego = p.est_tar = p.est_bl = p.est_tv = mode
            for _ in range(20):
                s, v, a = rng.uniform(0, 3000), rng.uniform(0, 30), rng.uniform(-3, 2)
                if rng.random() < 0.3:
                    v = rng.uniform(0, 0.5)                      # braking to (and holding at) standstill
                sp = np.cumsum(rng.uniform(0, 15, N + 1)) + s
                vp = rng.uniform(0, 30, N + 1)
                out = [np.zeros(N + 1) for _ in range(4)]
                L.k3r_emu_est_traj(C.byref(p), setting, s, v, a, sp.ctypes.data_as(dp), vp.ctypes.data_as(dp),
                                   out[0].ctypes.data_as(dp), out[1].ctypes.data_as(dp))
                O.orc_est_traj(C.byref(p), setting, s, v, a, sp.ctypes.data_as(dp), vp.ctypes.data_as(dp),
                               out[2].ctypes.data_as(dp), out[3].ctypes.data_as(dp))
                assert np.array_equal(out[0], out[2]) and np.array_equal(out[1], out[3]), (mode, setting)
                checked += 1
    assert checked == 240
    p = S.default_params("mc")
    for num in (3, 5, 6, 7, 11, 12):
        uc = U.get_use_case(num, S.default_params("mc"))
        r = uc["route"]
        for mpctype in (0, 1, 2):
            for _ in range(25):
                s_est = np.sort(rng.uniform(0, 1200, N + 1))
                v_est = rng.uniform(0, 35, N + 1)
                t0 = float(rng.uniform(0, 120))
                mine = np.zeros((N, 9))
                L.k3r_emu_bounds(C.byref(p), C.byref(r), mpctype, s_est.ctypes.data_as(dp), v_est.ctypes.data_as(dp), t0,
                                 mine.ctypes.data_as(dp))
                ref = [np.zeros(N) for _ in range(9)]
                O.orc_bounds(C.byref(p), C.byref(r), mpctype, s_est.ctypes.data_as(dp), v_est.ctypes.data_as(dp), t0,
                             *[a_.ctypes.data_as(dp) for a_ in ref])
                assert np.array_equal(mine, np.stack(ref, 1)), (num, mpctype)


@pytest.mark.parametrize("cn,first,n_scen,ns", [("fb", 0, 6, 160), ("ab", 3, 4, 160), ("fb", 41, 2, 871)])
def test_emulated_closed_loop_of_the_device_code_follows_the_oracle(oracle, cn, first, n_scen, ns):
    """The product's whole per-step device code (K2 build -> K3r reduced-space QP, a solve that outgrows the capacity
    repeated at capacity n -> K4 post), compiled for the host and run single-threaded, against the oracle on scenarios
    of the bench batch: north_star's bars, the same solver status, the same number of working-set changes and the
    same working set after EVERY solve -- on the CPU, standstill degeneracy and the cold-start step included.  (The
    GPU suite holds the CUDA build to the same bars; this test keeps the decision rules of csrc/qp_rules.cuh and
    oracle/qp_oracle.c:rt_select from drifting apart when no GPU is at hand.)"""
    import bench
    from eepacc_mpc_casadi_matlab_b200 import settings as S
    if R.build() is None:
        pytest.skip("no CUDA headers")
    cid = S.CTRL_NAMES[cn]
    p, routes, lead = bench.make_workload(cn, n_scen, first, ns)
    for b in range(n_scen):
        log, ws = R.emu_run(cid, p, routes[b], lead[:, b, :], fc=56, want_ws=True)
        V = oracle.Vehicle(cid, p, routes[b])
        for k in range(ns):
            o = V.step(lead[k, b, 0], lead[k, b, 1])
            g = log[k]
            assert g[9] == o[9] and g[10] == o[10], (b, k, g[9:11], o[9:11])
            assert np.max(np.abs(g[:2] - o[:2])) < 1e-6 and np.max(np.abs(g[2:4] - o[2:4])) < 1e-4, (b, k)
            assert abs(g[11] - o[11]) <= 1e-9 * (1 + abs(o[11])), (b, k, g[11], o[11])
            if k % 8 == 0 or k < 3:
                assert np.array_equal(ws[k], V.working_set()), (b, k)
