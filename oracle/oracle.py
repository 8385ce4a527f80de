"""ctypes binding of the CPU ORACLE (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY: import from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from eepacc_mpc_casadi_matlab_b200.settings import N_KPI, Params, Route

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_dp = C.POINTER(C.c_double)
NMAX, MMAX, NZMAX = 120, 522, 162


class OrcQP(C.Structure):
    _fields_ = [("n", C.c_int), ("m", C.c_int),
                ("H", C.c_double * (NMAX * NMAX)), ("g", C.c_double * NMAX),
                ("A", C.c_double * (MMAX * NMAX)), ("lba", C.c_double * MMAX),
                ("uba", C.c_double * MMAX), ("Psi", C.c_double * (NZMAX * NMAX)),
                ("d", C.c_double * NZMAX), ("nz", C.c_int)]


def build(force=False):
    if force or not os.path.exists(_LIB) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB)
            for f in ("eepacc_oracle.c", "qp_oracle.c", "nlp_oracle.c", "eepacc_oracle.h")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


def build_fast():
    """-O3 -march=native build of the same sources, for bench.py's CPU timing arm only (never for parity)."""
    fast = os.path.join(_HERE, "_build", "liboracle_fast.so")
    if not os.path.exists(fast) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(fast)
            for f in ("eepacc_oracle.c", "qp_oracle.c", "nlp_oracle.c", "eepacc_oracle.h")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "fast"])
    return fast


def run_batch_fast(ctrl, params, routes, lead):
    """Timing arm: B scenarios in one C call of the -O3 -march=native build.  lead [n_steps, B, 2] -> kpi [B, N_KPI]."""
    L = C.CDLL(build_fast())
    lead = np.ascontiguousarray(lead, dtype=np.float64)
    n, B = lead.shape[0], lead.shape[1]
    kpi = np.zeros((B, N_KPI)); scratch = np.zeros((n, 12))
    L.orc_run_batch.argtypes = [C.c_int, C.POINTER(Params), C.POINTER(Route), _dp, C.c_longlong, C.c_int, _dp, _dp]
    rc = L.orc_run_batch(ctrl, C.byref(params), routes, _p(lead), B, n, _p(kpi), _p(scratch))
    assert rc == 0, rc
    return kpi


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        L.orc_interp_pwa.restype = C.c_double
        L.orc_interp_pwa.argtypes = [C.c_double, _dp, _dp, C.c_int]
        L.orc_power5.restype = C.c_double
        L.orc_power5.argtypes = [_dp, C.c_double, C.c_double]
        L.orc_power2.restype = C.c_double
        L.orc_power2.argtypes = [_dp, C.c_double, C.c_double]
        L.orc_vehicle_new.restype = C.c_void_p
        L.orc_vehicle_new.argtypes = [C.c_int, C.POINTER(Params), C.POINTER(Route)]
        L.orc_vehicle_free.argtypes = [C.c_void_p]
        L.orc_vehicle_step.argtypes = [C.c_void_p, C.c_double, C.c_double, _dp]
        L.orc_vehicle_qp.restype = C.POINTER(OrcQP)
        L.orc_vehicle_qp.argtypes = [C.c_void_p]
        L.orc_vehicle_get_ws.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.orc_run.argtypes = [C.c_int, C.POINTER(Params), C.POINTER(Route), _dp, C.c_int, C.c_int,
                              _dp, _dp]
        L.orc_qpws_new.restype = C.c_void_p
        L.orc_vehicle_get_sol.argtypes = [C.c_void_p, _dp, _dp]
        L.orc_qpws_new.argtypes = [C.c_int, C.c_int]
        L.orc_qpws_free.argtypes = [C.c_void_p]
        L.orc_qpws_reset.argtypes = [C.c_void_p]
        L.orc_qpws_get_ws.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.orc_qpsolve.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp,
                                  C.POINTER(C.c_int)]
        L.orc_plant.argtypes = [C.POINTER(Params), C.POINTER(Route), _dp, _dp, _dp]
        L.orc_f.argtypes = [C.POINTER(Params), C.c_double, _dp, _dp, _dp, _dp]
        L.orc_F_jac.argtypes = [C.POINTER(Params), C.c_double, _dp, _dp, _dp, _dp, _dp]
        L.orc_qpbuild.argtypes = [C.c_int, C.POINTER(Params), C.POINTER(Route)] + [C.c_double] * 7 + \
                                 [_dp, _dp, C.c_void_p, C.c_int, C.POINTER(OrcQP)]
        L.orc_fb_ad_init.argtypes = [C.POINTER(Params), C.c_double, C.c_void_p]
        L.orc_kpi.argtypes = [C.POINTER(Params), C.c_double, _dp, C.c_int, _dp]
        L.orc_force_alloc.argtypes = [C.POINTER(Params), C.POINTER(Route)] + [C.c_double] * 3 + [_dp] * 3
        L.orc_lut_linear.restype = C.c_double
        L.orc_lut_linear.argtypes = [C.c_double, _dp, _dp, C.c_int]
        L.orc_nlp_rows_per_stage.argtypes = [C.POINTER(Route)]
        L.orc_nlp_eval.argtypes = [C.POINTER(Params), C.POINTER(Route), _dp, _dp, C.c_int, _dp, _dp, C.c_int, _dp, C.c_int,
                                   _dp] + [C.c_void_p] * 6
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def qp_arrays(qp):
    """dense (H, g, A, lba, uba) of an OrcQP as numpy (H n x n, A m x n)."""
    n, m = qp.n, qp.m
    H = np.frombuffer(qp.H, dtype=np.float64, count=n * n).reshape(n, n).T.copy()
    A = np.frombuffer(qp.A, dtype=np.float64, count=n * m).reshape(n, m).T.copy()
    g = np.frombuffer(qp.g, dtype=np.float64, count=n).copy()
    lba = np.frombuffer(qp.lba, dtype=np.float64, count=m).copy()
    uba = np.frombuffer(qp.uba, dtype=np.float64, count=m).copy()
    return H, g, A, lba, uba


class Vehicle:
    """One closed-loop vehicle (orc_vehicle_step == one RunOpt_*MPC.m loop iteration)."""

    def __init__(self, ctrl, params, route):
        self.L = lib()
        self.params, self.route = params, route
        self.h = self.L.orc_vehicle_new(ctrl, C.byref(params), C.byref(route))
        self.ctrl = ctrl

    def step(self, s_tv, v_tv):
        out = np.zeros(12)
        self.L.orc_vehicle_step(self.h, float(s_tv), float(v_tv), _p(out))
        return out

    def qp(self):
        return self.L.orc_vehicle_qp(self.h).contents

    def solution(self):
        """(x, lam_a) of the last closed-loop QP."""
        q = self.qp()
        x, lam = np.zeros(q.n), np.zeros(q.m)
        self.L.orc_vehicle_get_sol(self.h, _p(x), _p(lam))
        return x, lam

    def working_set(self):
        q = self.qp()
        ws = (C.c_int * (q.n + q.m))()
        self.L.orc_vehicle_get_ws(self.h, ws)
        return np.array(ws[:], dtype=np.int32)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_vehicle_free(self.h)
            self.h = None


def run(ctrl, params, route, lead):
    """lead: float64[n_steps, 2] -> (log[n_steps, 12], kpi[N_KPI])."""
    lead = np.ascontiguousarray(lead, dtype=np.float64)
    n = lead.shape[0]
    log = np.zeros((n, 12))
    kpi = np.zeros(N_KPI)
    lib().orc_run(ctrl, C.byref(params), C.byref(route), _p(lead), 2, n, _p(log), _p(kpi))
    return log, kpi


class QPSolver:
    """Warm-started dense QP solver object (the conic('qpoases') Function of the reference)."""

    def __init__(self, n, m):
        self.L = lib()
        self.n, self.m = n, m
        self.h = self.L.orc_qpws_new(n, m)

    def reset(self):
        self.L.orc_qpws_reset(self.h)

    def solve(self, H, g, A, lba, uba, lbx=None, ubx=None):
        n, m = self.n, self.m
        Hc = np.asfortranarray(H, dtype=np.float64)
        Ac = np.asfortranarray(np.asarray(A, dtype=np.float64).reshape(m, n))
        g = np.ascontiguousarray(g, dtype=np.float64)
        lba = np.ascontiguousarray(lba, dtype=np.float64)
        uba = np.ascontiguousarray(uba, dtype=np.float64)
        x, lam_a, lam_x, cost = np.zeros(n), np.zeros(m), np.zeros(n), np.zeros(1)
        it = C.c_int(0)
        lbx_p = _p(np.ascontiguousarray(lbx, dtype=np.float64)) if lbx is not None else None
        ubx_p = _p(np.ascontiguousarray(ubx, dtype=np.float64)) if ubx is not None else None
        st = self.L.orc_qpsolve(self.h, Hc.ctypes.data_as(_dp), _p(g), Ac.ctypes.data_as(_dp),
                                _p(lba), _p(uba), lbx_p, ubx_p, _p(x), _p(cost), _p(lam_a),
                                _p(lam_x), C.byref(it))
        return dict(x=x, cost=cost[0], lam_a=lam_a, lam_x=lam_x, status=st, iter=it.value)

    def working_set(self):
        ws = (C.c_int * (self.n + self.m))()
        self.L.orc_qpws_get_ws(self.h, ws)
        return np.array(ws[:], dtype=np.int32)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_qpws_free(self.h)
            self.h = None


def nlp_eval(params, route, s_velInc, v_velInc, s_stop, v_stop, s_tv, z):
    """RunOpt_NLP.m's problem functions at z (reference layout): dict with f, g, lbg, ubg, lbz, ubz (nlp_oracle.c)."""
    L = lib()
    z = np.ascontiguousarray(z, dtype=np.float64)
    N = (len(z) - 4) // 10
    s_tv = np.ascontiguousarray(s_tv[:N + 1], dtype=np.float64)
    nr = N * L.orc_nlp_rows_per_stage(C.byref(route))
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (s_velInc, v_velInc, s_stop, v_stop)]
    f = np.zeros(1); g = np.zeros(nr); lbg = np.zeros(nr); ubg = np.zeros(nr); lbz = np.zeros(len(z)); ubz = np.zeros(len(z))
    n = L.orc_nlp_eval(C.byref(params), C.byref(route), _p(a[0]), _p(a[1]), len(a[0]), _p(a[2]), _p(a[3]), len(a[2]),
                       _p(s_tv), N, _p(z), f.ctypes.data, g.ctypes.data, lbg.ctypes.data, ubg.ctypes.data,
                       lbz.ctypes.data, ubz.ctypes.data)
    assert n == nr, (n, nr)
    return dict(f=float(f[0]), g=g, lbg=lbg, ubg=ubg, lbz=lbz, ubz=ubz)


def nlp_violation(ev, z):
    """Largest violation of lbg <= g <= ubg and lbz <= z <= ubz."""
    vg = np.maximum(ev["lbg"] - ev["g"], ev["g"] - ev["ubg"]).max()
    vz = np.maximum(ev["lbz"] - z, z - ev["ubz"]).max()
    return float(max(vg, vz, 0.0))
