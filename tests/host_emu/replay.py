"""TEST INFRASTRUCTURE: replay recorded closed-loop QP operands through the host emulation of k3r_qp.cu."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from eepacc_mpc_casadi_matlab_b200 import settings as S  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class QpIO(C.Structure):
    _fields_ = [("B", C.c_longlong), ("n", C.c_int), ("m", C.c_int), ("ctrl", C.c_int),
                ("H", _dp), ("g", _dp), ("A", _dp), ("lba", _dp), ("uba", _dp), ("lbx", _dp), ("ubx", _dp),
                ("sblk", _dp), ("wx", _dp), ("wy", _dp), ("widx", C.POINTER(C.c_short)),
                ("wside", C.POINTER(C.c_byte)), ("wnA", _ip), ("warm", _ip),
                ("x", _dp), ("cost", _dp), ("lam_a", _dp), ("lam_x", _dp), ("status", _dp), ("iter", _dp),
                ("counters", C.POINTER(C.c_ulonglong)), ("dbg", _dp), ("dbg_cap", C.c_int), ("fc", C.c_int),
                ("defer_list", _ip), ("defer_cnt", _ip), ("todo_list", _ip), ("todo_cnt", _ip),
                ("rowtab", C.c_void_p), ("rowc", _dp), ("hscr", _dp), ("todo_grid", C.c_int)]


def cuda_include():
    for d in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if d and os.path.exists(os.path.join(d, "include", "cuda_runtime.h")):
            return os.path.join(d, "include")
    return None


def build(opt="-O2"):
    inc = cuda_include()
    if inc is None:
        return None
    out = os.path.join(HERE, "_build")
    os.makedirs(out, exist_ok=True)
    lib = os.path.join(out, "libk3r_emu.so")
    src = os.path.join(HERE, "k3r_emu.cpp")
    deps = [src] + [os.path.join(ROOT, "eepacc_mpc_casadi_matlab_b200", "csrc", f)
                    for f in ("k3r_qp.cu", "eepacc_rows.cuh", "eepacc_dev.cuh", "eepacc_kernels.h", "k2_build.cuh", "k4_post.cuh",
                              "k6_ocp.cuh", "qp_rules.cuh")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.check_call(["g++", opt, "-g", "-std=c++17", "-x", "c++", "-shared", "-fPIC", "-I", inc, src, "-o", lib])
    L = C.CDLL(lib)
    L.k3r_emu_sizeof_io.restype = C.c_longlong
    assert L.k3r_emu_sizeof_io() == C.sizeof(QpIO), (L.k3r_emu_sizeof_io(), C.sizeof(QpIO))
    return L


class Replayer:
    """Warm-started solver state for B vehicles, fed one recorded step at a time."""

    def __init__(self, L, ctrl, n, m, B, fc, params=None):
        self.L, self.ctrl, self.n, self.m, self.B, self.fc = L, ctrl, n, m, B, fc
        self.p = params if params is not None else S.default_params("mc")
        self.wx = np.zeros((B, n)); self.wy = np.zeros((B, n))
        self.widx = np.zeros((B, n), np.int16); self.wside = np.zeros((B, n), np.int8)
        self.wnA = np.zeros(B, np.int32); self.warm = np.zeros(B, np.int32)
        self.x = np.zeros((B, n)); self.cost = np.zeros(B); self.status = np.zeros(B); self.iter = np.zeros(B)
        self.counters = np.zeros((B, 4), np.uint64)
        self.defer_list = np.zeros(B, np.int32); self.defer_cnt = np.zeros(1, np.int32)
        self.dbg = np.zeros(4096 * 4 + 32)

    def solve(self, sblk, g, lba, uba):
        io = QpIO()
        io.B, io.n, io.m, io.ctrl, io.fc = self.B, self.n, self.m, self.ctrl, self.fc
        keep = [np.ascontiguousarray(a, dtype=np.float64) for a in (sblk, g, lba, uba)]
        io.sblk, io.g, io.lba, io.uba = [a.ctypes.data_as(_dp) for a in keep]
        io.wx = self.wx.ctypes.data_as(_dp); io.wy = self.wy.ctypes.data_as(_dp)
        io.widx = self.widx.ctypes.data_as(C.POINTER(C.c_short))
        io.wside = self.wside.ctypes.data_as(C.POINTER(C.c_byte))
        io.wnA = self.wnA.ctypes.data_as(_ip); io.warm = self.warm.ctypes.data_as(_ip)
        io.x = self.x.ctypes.data_as(_dp); io.cost = self.cost.ctypes.data_as(_dp)
        io.status = self.status.ctypes.data_as(_dp); io.iter = self.iter.ctypes.data_as(_dp)
        io.counters = self.counters.ctypes.data_as(C.POINTER(C.c_ulonglong))
        io.dbg = self.dbg.ctypes.data_as(_dp); io.dbg_cap = 4096
        self.defer_cnt[0] = 0
        io.defer_list = self.defer_list.ctypes.data_as(_ip); io.defer_cnt = self.defer_cnt.ctypes.data_as(_ip)
        rc = self.L.k3r_emu_solve(C.byref(self.p), C.byref(io))
        assert rc == 0
        dl = self.defer_list[:self.defer_cnt[0]].copy()
        # deferred vehicles: what the full-space kernel does on the GPU is emulated here by the same code
        # with the capacity set to n (it then never defers)
        for b in dl:
            sub = Replayer(self.L, self.ctrl, self.n, self.m, 1, self.n, self.p)
            names = ("wx", "wy", "widx", "wside", "wnA", "warm", "counters")
            for nm in names:
                getattr(sub, nm)[0] = getattr(self, nm)[b]
            assert len(sub.solve(keep[0][b:b + 1], keep[1][b:b + 1], keep[2][b:b + 1], keep[3][b:b + 1])) == 0
            for nm in names + ("x", "cost", "status", "iter"):
                getattr(self, nm)[b] = getattr(sub, nm)[0]
        return dl


def replay(path, fc=None, steps=None, verbose=True):
    d = np.load(path)
    n, m = int(d["n"]), int(d["m"])
    ctrl = {"ab": S.AB, "fb": S.FB, "bl": S.BL, "tv": S.TV}[os.path.basename(path).split("_")[1].split(".")[0]]
    ns, B = d["x"].shape[:2]
    if steps:
        ns = min(ns, steps)
    L = build()
    R = Replayer(L, ctrl, n, m, B, fc or n)
    worst = 0.0
    for k in range(ns):
        dl = R.solve(d["sblk"][k], d["g"][k], d["lba"][k], d["uba"][k])
        ex = np.abs(R.x - d["x"][k]).max(axis=1)
        ec = np.abs(R.cost - d["cost"][k, :, 0]) / (1 + np.abs(d["cost"][k, :, 0]))
        bad = (R.status != d["status"][k, :, 0])
        worst = max(worst, ex.max())
        if verbose and (ex.max() > 1e-6 or bad.any() or len(dl) or k % 50 == 0):
            print(f"step {k}: max|dx| {ex.max():.3e} cost rel {ec.max():.2e} status {R.status.astype(int)} ref {d['status'][k,:,0].astype(int)}"
                  f" iter {R.iter.astype(int)} ref {d['iter'][k,:,0].astype(int)} defer {dl}")
    return worst


if __name__ == "__main__":
    w = replay(sys.argv[1], fc=int(sys.argv[2]) if len(sys.argv) > 2 else None,
               steps=int(sys.argv[3]) if len(sys.argv) > 3 else None)
    print("worst |x - x_ref| =", w)


def emu_run(ctrl, params, route, lead, fc=56, want_ws=False, stats=None):
    """Closed loop of ONE vehicle through the host emulation of the product's device code (K2 -> K3r -> K4).
    lead: float64[n_steps, 2] (None for TVMPC with n_steps given as an int).  -> log[n_steps, 12] (, ws[n_steps, m+n])."""
    from eepacc_mpc_casadi_matlab_b200.settings import Params, Route
    L = build()
    L.k3r_emu_run.argtypes = [C.c_int, C.POINTER(Params), C.POINTER(Route), _dp, C.c_int, C.c_int, _dp, C.c_void_p, _ip]
    if isinstance(lead, int):
        ns, lp = lead, None
    else:
        lead = np.ascontiguousarray(lead, dtype=np.float64)
        ns, lp = lead.shape[0], lead.ctypes.data_as(_dp)
    log = np.zeros((ns, 12))
    n = {S.FB: 120, S.AB: 100}.get(ctrl, 40)
    m = {S.FB: 522, S.AB: 362 if params.ab_route_rows else 282, S.BL: 262, S.TV: 220}[ctrl]
    ws = np.zeros((ns, n + m), np.int8) if want_ws else None
    rc = L.k3r_emu_run(ctrl, C.byref(params), C.byref(route), lp, ns, fc, log.ctypes.data_as(_dp),
                       ws.ctypes.data_as(C.c_void_p) if want_ws else None,
                       stats.ctypes.data_as(_ip) if stats is not None else None)
    assert rc == 0, rc
    return (log, ws) if want_ws else log


OCP_DEFAULTS = dict(mu0=1e6, eps_mu=1e-11, eps_p=1e-9, eps_d=1e-10, maxit=80, warm=0, eps_accept=1e-9, gamma=1e-3, tau0=1.0)   # = the library defaults (eepacc_capi.cu: ocp_opts)


def ocp_emu_run(ctrl, params, route, lead, check=True, **opts):
    """Closed loop of ONE vehicle with the stage-structured interior-point solver K6 (csrc/k6_ocp.cuh) in the
    solver's place, single-threaded on the host: K2 -> K6 -> K4.  -> (log[n_steps, 12], diag[n_steps, 6]);
    diag = (first-stage / all-stage max |u - u_dense|, relative objective difference, K6 iterations, K6 status,
    dense status) against the reduced-space active-set kernel on the same QP operands when check is set."""
    from eepacc_mpc_casadi_matlab_b200.settings import Params, Route
    L = build()
    L.k6_emu_run.argtypes = [C.c_int, C.POINTER(Params), C.POINTER(Route), _dp, C.c_int, _dp, _dp, _dp]
    lead = np.ascontiguousarray(lead, dtype=np.float64)
    ns = lead.shape[0]
    o = dict(OCP_DEFAULTS); o["mu0"] = 1e4 if ctrl == S.FB else 1e6; o.update(opts)       # eepacc_capi.cu: eepacc_mem_set_solver
    ov = np.array([o["mu0"], o["eps_mu"], o["eps_p"], o["eps_d"], o["maxit"], o["warm"], 1.0 if check else 0.0, o["eps_accept"], o["gamma"], o["tau0"]])
    log = np.zeros((ns, 12)); diag = np.zeros((ns, 6))
    rc = L.k6_emu_run(ctrl, C.byref(params), C.byref(route), lead.ctypes.data_as(_dp), ns, ov.ctypes.data_as(_dp),
                      log.ctypes.data_as(_dp), diag.ctypes.data_as(_dp))
    assert rc == 0, rc
    return log, diag
