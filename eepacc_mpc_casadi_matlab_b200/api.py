"""Host-side mirror of the reference's operator interface, over the C-ABI of libeepacc.so.

Reference                                        here
---------------------------------------------    -----------------------------------------
optSol = RunOpt_FBMPC(OPTsettings)               RunOpt_FBMPC(params, routes, lead) -> optSol dict
  (RunOpt_FBMPC.m:1; also AB/BL/TV)                (same field names: s_opt, v_opt, Fm_opt, ...)
[H,c,G,..] = CreateQP_FB(..) + TransformTo..     Engine.qpbuild(x0, a_minus1, lead, t0)
sol = QPsolver('h',H,'g',c,'a',G,..)             Engine.qpsolve(h, g, a, lba, uba, lbx, ubx)
RunPlantModel(x, u, OPTsettings)                 Engine.plant(x, u)
f(x,u), F(x0,p) CasADi Functions (RunOpt_NLP)    Engine.f / Engine.F / Engine.F_jac
GetMotorPower_FifthOrderSurface / QuadraticFit   Engine.power5 / Engine.power2

All arrays carry a leading batch dimension.  numpy arrays are host buffers (copied inside the
call); torch CUDA tensors are passed as device pointers (no copies).  There is no CPU fallback:
if libeepacc.so is missing or no CUDA device is usable the call raises.
"""
import ctypes as C
import os

import numpy as np

from . import settings as S

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EEPACC_LIB") or os.path.join(_HERE, "libeepacc.so")   # EEPACC_LIB: developer variant (timers)
_dp = C.POINTER(C.c_double)
_dpp = C.POINTER(_dp)

FAMILIES = ["eepacc_f", "eepacc_F", "eepacc_F_jac", "eepacc_power5", "eepacc_power2", "eepacc_plant",
            "eepacc_qpbuild_ab", "eepacc_qpbuild_fb", "eepacc_qpbuild_bl", "eepacc_qpbuild_tv",
            "eepacc_qpsolve", "eepacc_step_ab", "eepacc_step_fb", "eepacc_step_bl", "eepacc_step_tv",
            "eepacc_run"]
FAMILY_SUFFIXES = ["", "_n_in", "_n_out", "_name_in", "_name_out", "_sparsity_in", "_sparsity_out",
                   "_work", "_incref", "_decref", "_checkout", "_release", "_default_in"]
PLAIN_SYMBOLS = ["eepacc_qp_nvar", "eepacc_qp_ncon", "eepacc_mem_setup", "eepacc_mem_reset",
                 "eepacc_mem_set_step", "eepacc_mem_set_lead", "eepacc_mem_set_stream",
                 "eepacc_mem_launch_count", "eepacc_mem_get_counters", "eepacc_mem_kpi_device_ptr",
                 "eepacc_sizeof_params", "eepacc_sizeof_route", "eepacc_last_error",
                 "eepacc_measure_dfma_peak", "eepacc_mem_debug_trace", "eepacc_mem_enable_timing",
                 "eepacc_mem_get_timing", "eepacc_mem_debug_phases"]


class EepaccError(RuntimeError):
    pass


_lib = None


# eepacc_run pipelines up to 16 slices on their own streams; with the default 8 hardware queues slices share a queue
# and serialise.  Only effective if CUDA has not been initialised yet in this process.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def load_library():
    """Load libeepacc.so (built in-tree by eepacc_mpc_casadi_matlab_b200.build).  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EepaccError(f"{LIB_PATH} is missing: build it with "
                          "`python -m eepacc_mpc_casadi_matlab_b200.build` (no CPU fallback exists)")
    L = C.CDLL(LIB_PATH)
    L.eepacc_last_error.restype = C.c_char_p
    L.eepacc_mem_setup.argtypes = [C.c_int, C.c_int, C.POINTER(S.Params), C.POINTER(S.Route), C.c_longlong, C.c_int]
    L.eepacc_mem_set_lead.argtypes = [C.c_int, _dp, C.c_longlong]
    L.eepacc_mem_synth_lead.argtypes = [C.c_int, _dp, C.c_longlong, _dp, C.c_longlong, _dp, C.c_longlong, C.c_double]
    L.eepacc_mem_set_step.argtypes = [C.c_int, C.c_int, C.c_double]
    L.eepacc_mem_gen_routes.argtypes = [C.c_int, C.c_void_p, C.c_longlong]
    L.eepacc_mem_fetch_routes.argtypes = [C.c_int, C.c_void_p, C.c_longlong]
    L.eepacc_sizeof_route_spec.restype = C.c_longlong
    L.eepacc_mem_set_stream.argtypes = [C.c_int, C.c_void_p]
    L.eepacc_mem_launch_count.restype = C.c_longlong
    L.eepacc_mem_get_counters.argtypes = [C.c_int, _dp]
    L.eepacc_mem_kpi_device_ptr.restype = C.c_void_p
    L.eepacc_sizeof_params.restype = C.c_longlong
    L.eepacc_sizeof_route.restype = C.c_longlong
    L.eepacc_measure_dfma_peak.restype = C.c_double
    L.eepacc_measure_dfma_peak.argtypes = [C.c_int]
    L.eepacc_qp_ncon.argtypes = [C.c_int, C.POINTER(S.Params)]
    L.eepacc_mem_debug_trace.argtypes = [C.c_int, _dp, C.c_int]
    L.eepacc_mem_enable_timing.argtypes = [C.c_int, C.c_int]
    L.eepacc_mem_debug_phases.argtypes = [C.c_int, _dp]
    L.eepacc_mem_get_timing.argtypes = [C.c_int, _dp, _dp, C.c_longlong]
    L.eepacc_mem_set_solver.argtypes = [C.c_int, C.c_int]
    L.eepacc_mem_get_solver.argtypes = [C.c_int]
    L.eepacc_mem_set_ocp_options.argtypes = [C.c_int, _dp]
    for f in FAMILIES:
        getattr(L, f).argtypes = [_dpp, _dpp, C.POINTER(C.c_longlong), _dp, C.c_int]
        getattr(L, f + "_name_in").restype = C.c_char_p
        getattr(L, f + "_name_in").argtypes = [C.c_longlong]
        getattr(L, f + "_name_out").restype = C.c_char_p
        getattr(L, f + "_name_out").argtypes = [C.c_longlong]
        getattr(L, f + "_sparsity_in").restype = C.POINTER(C.c_longlong)
        getattr(L, f + "_sparsity_in").argtypes = [C.c_longlong]
        getattr(L, f + "_sparsity_out").restype = C.POINTER(C.c_longlong)
        getattr(L, f + "_sparsity_out").argtypes = [C.c_longlong]
        getattr(L, f + "_n_in").restype = C.c_longlong
        getattr(L, f + "_n_out").restype = C.c_longlong
        getattr(L, f + "_release").argtypes = [C.c_int]
    if (L.eepacc_sizeof_params() != C.sizeof(S.Params) or L.eepacc_sizeof_route() != C.sizeof(S.Route)
            or L.eepacc_sizeof_route_spec() != C.sizeof(S.RouteSpec)):
        raise EepaccError("ABI mismatch between settings.py and include/eepacc.h")
    _lib = L
    return L


def _is_device(a):
    return hasattr(a, "data_ptr") and getattr(a, "is_cuda", False)


def _ptr(a):
    if a is None:
        return C.cast(None, _dp)
    if _is_device(a):
        return C.cast(C.c_void_p(a.data_ptr()), _dp)
    return a.ctypes.data_as(_dp)


def _check_buffer(a, device_index):
    """The C-ABI takes raw double pointers: refuse anything that is not contiguous float64 (on the engine's GPU)."""
    if a is None:
        return
    if _is_device(a):
        import torch
        if a.dtype != torch.float64:
            raise EepaccError(f"device buffer has dtype {a.dtype}, the ABI computes in float64")
        if not a.is_contiguous():
            raise EepaccError("device buffer is not contiguous")
        if device_index is not None and a.device.index is not None and a.device.index != device_index:
            raise EepaccError(f"device buffer lives on cuda:{a.device.index}, the engine on cuda:{device_index}")
    else:
        if a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
            raise EepaccError("host buffer must be C-contiguous float64")


def _host(a, shape=None):
    if a is None or _is_device(a):
        return a
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def routes_array(routes, B):
    """Accept one Route (replicated), a list, or a ctypes array of B routes."""
    if isinstance(routes, S.Route):
        arr = (S.Route * B)()
        for b in range(B):
            C.memmove(C.byref(arr[b]), C.byref(routes), C.sizeof(S.Route))
        return arr
    if isinstance(routes, (list, tuple)):
        arr = (S.Route * B)()
        for b in range(B):
            C.memmove(C.byref(arr[b]), C.byref(routes[b]), C.sizeof(S.Route))
        return arr
    return routes


class Engine:
    """One `mem` object of the C-ABI: B vehicles of one controller on one GPU."""

    def __init__(self, ctrl, params, routes, B=None, device=-1):
        self.L = load_library()
        if isinstance(ctrl, str):
            ctrl = S.CTRL_NAMES[ctrl.lower()]
        self.ctrl = ctrl
        self.params = params
        if B is None:
            B = 1 if isinstance(routes, S.Route) else len(routes)
        self.B = int(B)
        # routes=None: the route blocks are built on the device afterwards (gen_routes) -- nothing is uploaded
        self.routes = routes_array(routes, self.B) if routes is not None else None
        self._fam = ["eepacc_step_ab", "eepacc_step_fb", "eepacc_step_bl", "eepacc_step_tv"][ctrl]
        self.mem = getattr(self.L, self._fam + "_checkout")()
        if self.mem < 0:
            raise EepaccError(self._err())
        rc = self.L.eepacc_mem_setup(self.mem, ctrl, C.byref(params), self.routes, self.B, device)     # routes None -> NULL
        if rc:
            msg = self._err()
            self.close()
            raise EepaccError(f"eepacc_mem_setup failed ({rc}): {msg}")
        self.n = self.L.eepacc_qp_nvar(ctrl)
        self.m = self.L.eepacc_qp_ncon(ctrl, C.byref(params))
        self.L.eepacc_mem_device.argtypes = [C.c_int]
        self.device_index = int(self.L.eepacc_mem_device(self.mem))

    # ---- plumbing ----
    def _err(self):
        e = self.L.eepacc_last_error()
        return e.decode() if e else ""

    def close(self):
        if getattr(self, "mem", -1) is not None and getattr(self, "mem", -1) >= 0:
            getattr(self.L, self._fam + "_release")(self.mem)
            self.mem = -1

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _call(self, name, args, ress, B):
        dev = any(_is_device(a) for a in list(args) + list(ress) if a is not None)
        if dev and any((a is not None and not _is_device(a)) for a in list(args) + list(ress)):
            raise EepaccError("mixing host and device buffers in one call is not supported")
        for a in list(args) + list(ress):
            _check_buffer(a, self.device_index)
        argv = (_dp * max(1, len(args)))(*[_ptr(a) for a in args])
        resv = (_dp * max(1, len(ress)))(*[_ptr(a) for a in ress])
        iw = (C.c_longlong * 2)(int(B), S.FLAG_DEVICE_PTR if dev else 0)
        rc = getattr(self.L, name)(argv, resv, iw, None, self.mem)
        if rc:
            raise EepaccError(f"{name} failed ({rc}): {self._err()}")

    def _out(self, like, shape):
        if _is_device(like):
            import torch
            return torch.empty(shape, dtype=torch.float64, device=like.device)
        return np.empty(shape, dtype=np.float64)

    # ---- K1 / K4 ----
    def f(self, x, u):
        x, u = _host(x), _host(u)
        B = x.shape[0]
        xdot, Lc = self._out(x, (B, 4)), self._out(x, (B,))
        self._call("eepacc_f", [x, u], [xdot, Lc], B)
        return xdot, Lc

    def F(self, x0, p):
        x0, p = _host(x0), _host(p)
        B = x0.shape[0]
        xf, qf = self._out(x0, (B, 4)), self._out(x0, (B,))
        self._call("eepacc_F", [x0, p], [xf, qf], B)
        return xf, qf

    def F_jac(self, x0, p):
        """Returns xf [B,4], qf [B], J [B,5,10] (J[b] = d[xf;qf]/d[x0;p])."""
        x0, p = _host(x0), _host(p)
        B = x0.shape[0]
        xf, qf, J = self._out(x0, (B, 4)), self._out(x0, (B,)), self._out(x0, (B, 10, 5))
        self._call("eepacc_F_jac", [x0, p], [xf, qf, J], B)
        return xf, qf, (J.transpose(1, 2) if _is_device(J) else J.transpose(0, 2, 1))

    def power5(self, Fm, rpm):
        Fm, rpm = _host(Fm), _host(rpm)
        P = self._out(Fm, Fm.shape)
        self._call("eepacc_power5", [Fm, rpm], [P], int(np.prod(Fm.shape)))
        return P

    def power2(self, Fm, rpm):
        Fm, rpm = _host(Fm), _host(rpm)
        P = self._out(Fm, Fm.shape)
        self._call("eepacc_power2", [Fm, rpm], [P], int(np.prod(Fm.shape)))
        return P

    def plant(self, x, u):
        x, u = _host(x), _host(u)
        B = x.shape[0]
        xn = self._out(x, (B, 2))
        self._call("eepacc_plant", [x, u], [xn], B)
        return xn

    # ---- K5r: route previews built on the device ----
    def gen_routes(self, specs):
        """specs: usecase.SPEC_DTYPE[B] (spec_table_fast), a RouteSpec ctypes array, or a list of RouteSpec:
        k5_route_gen_kernel builds the B route blocks in HBM (GenerateUseCase.m + SimplifyPWA.m per scenario) and
        the closed-loop state is reset."""
        if isinstance(specs, np.ndarray):
            specs = np.ascontiguousarray(specs)
            if specs.shape != (self.B,) or specs.dtype.itemsize != C.sizeof(S.RouteSpec):
                raise EepaccError("specs must be B eepacc_route_spec records")
            ptr, keep = specs.ctypes.data, specs
        else:
            if isinstance(specs, (list, tuple)):
                arr = (S.RouteSpec * self.B)()
                for b in range(self.B):
                    C.memmove(C.byref(arr[b]), C.byref(specs[b]), C.sizeof(S.RouteSpec))
                specs = arr
            ptr, keep = C.addressof(specs), specs
        rc = self.L.eepacc_mem_gen_routes(self.mem, ptr, self.B)
        del keep
        if rc:
            raise EepaccError(f"eepacc_mem_gen_routes failed ({rc}): {self._err()}")

    def fetch_routes(self):
        """The mem's B route blocks as a ctypes Route array (tests: bit-compare with usecase.generate_use_case)."""
        out = (S.Route * self.B)()
        rc = self.L.eepacc_mem_fetch_routes(self.mem, C.addressof(out), self.B)
        if rc:
            raise EepaccError(f"eepacc_mem_fetch_routes failed ({rc}): {self._err()}")
        return out

    # ---- K2 / K3 ----
    def set_step(self, kk, t0):
        self.L.eepacc_mem_set_step(self.mem, int(kk), float(t0))

    def qpbuild(self, x0, a_minus1, lead, t0):
        """-> h [B,n,n], g [B,n], a [B,m,n], lba [B,m], uba [B,m] (conic scheme names)."""
        B, n, m = self.B, self.n, self.m
        x0, a_minus1, lead, t0 = _host(x0, (B, 2)), _host(a_minus1, (B,)), _host(lead, (B, 3)), _host(t0, (B,))
        h, g = self._out(x0, (B, n, n)), self._out(x0, (B, n))
        a, lba, uba = self._out(x0, (B, n, m)), self._out(x0, (B, m)), self._out(x0, (B, m))
        name = ["eepacc_qpbuild_ab", "eepacc_qpbuild_fb", "eepacc_qpbuild_bl", "eepacc_qpbuild_tv"][self.ctrl]
        self._call(name, [x0, a_minus1, lead, t0], [h, g, a, lba, uba], B)
        at = a.transpose(1, 2) if _is_device(a) else a.transpose(0, 2, 1)   # column-major [n][m] -> [m, n] view
        return h, g, at, lba, uba

    def qpsolve(self, h, g, a, lba, uba, lbx=None, ubx=None):
        """Dense QP in the conic scheme; `a` is [B,m,n] (row-major view) and is re-laid column-major."""
        B, n, m = self.B, self.n, self.m
        dev = _is_device(h)
        if dev:
            a_cm = a.transpose(1, 2).contiguous()
            h_cm = h.transpose(1, 2).contiguous()
        else:
            a_cm = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(B, m, n).transpose(0, 2, 1))
            h_cm = np.ascontiguousarray(np.asarray(h, dtype=np.float64).reshape(B, n, n).transpose(0, 2, 1))
        g, lba, uba = _host(g, (B, n)), _host(lba, (B, m)), _host(uba, (B, m))
        lbx = _host(lbx, (B, n)) if lbx is not None else None
        ubx = _host(ubx, (B, n)) if ubx is not None else None
        x, cost = self._out(g, (B, n)), self._out(g, (B,))
        lam_a, lam_x = self._out(g, (B, m)), self._out(g, (B, n))
        status, it = self._out(g, (B,)), self._out(g, (B,))
        self._call("eepacc_qpsolve", [h_cm, g, a_cm, lba, uba, lbx, ubx, None, None, None],
                   [x, cost, lam_a, lam_x, status, it], B)
        return dict(x=x, cost=cost, lam_a=lam_a, lam_x=lam_x, status=status, iter=it)

    # ---- closed loop ----
    def reset(self):
        rc = self.L.eepacc_mem_reset(self.mem)
        if rc:
            raise EepaccError(self._err())

    def step(self, lead=None):
        """One `for kk = 0:N_sim` iteration for all B vehicles.  lead: [B,2] = (s_tv(k), v_tv(k))."""
        B = self.B
        lead = _host(lead, (B, 2)) if lead is not None else None
        like = lead if lead is not None else np.empty(0)
        z1, a, st, it = self._out(like, (B, 8)), self._out(like, (B,)), self._out(like, (B,)), self._out(like, (B,))
        self._call(self._fam, [lead], [z1, a, st, it], B)
        return dict(z1=z1, a=a, status=st, iter=it)

    def set_lead(self, lead):
        """lead: host float64 [n_steps, B, 2]."""
        lead = np.ascontiguousarray(lead, dtype=np.float64)
        assert lead.shape[1] == self.B and lead.shape[2] == 2
        rc = self.L.eepacc_mem_set_lead(self.mem, lead.ctypes.data_as(_dp), lead.shape[0])
        if rc:
            raise EepaccError(self._err())
        self._n_lead = lead.shape[0]

    def synth_lead(self, pert, base_v_2hz, highway_v_2hz=None, n_steps=871, tv_length=4.0):
        """Make the lead traces of a perturbed scenario batch on the device (k5_lead_synth_kernel) instead of
        uploading them: pert float64[B, N_PERT] from usecase.perturbation_table, 2 Hz cycle samples as in
        usecase.synth_lead_host (its host statement)."""
        pert = np.ascontiguousarray(pert, dtype=np.float64)
        assert pert.shape == (self.B, S.N_PERT)
        base = np.ascontiguousarray(base_v_2hz, dtype=np.float64)
        hw = None if highway_v_2hz is None else np.ascontiguousarray(highway_v_2hz, dtype=np.float64)
        rc = self.L.eepacc_mem_synth_lead(self.mem, base.ctypes.data_as(_dp), base.size,
                                          hw.ctypes.data_as(_dp) if hw is not None else None,
                                          hw.size if hw is not None else 0, pert.ctypes.data_as(_dp),
                                          int(n_steps), float(tv_length))
        if rc:
            raise EepaccError(f"eepacc_mem_synth_lead failed ({rc}): {self._err()}")
        self._n_lead = int(n_steps)

    def fetch_lead(self):
        """The lead trace resident on the device (uploaded or synthesised) -> host [n_steps, B, 2]."""
        return self.debug_fetch(9, self._n_lead * self.B * 2).reshape(self._n_lead, self.B, 2)

    def run(self, n_steps, want_log=True, device_out=None):
        """Whole closed loop on the device (lead set by set_lead).  -> (log [n_steps,B,12] | None, kpi [B,8])."""
        ns = np.array([float(n_steps)])
        like = device_out if device_out is not None else ns
        log = self._out(like, (n_steps, self.B, 12)) if want_log else None
        kpi = self._out(like, (self.B, S.N_KPI))
        dev = _is_device(kpi)
        argv = (_dp * 1)(ns.ctypes.data_as(_dp))
        resv = (_dp * 2)(_ptr(log), _ptr(kpi))
        iw = (C.c_longlong * 2)(self.B, S.FLAG_DEVICE_PTR if dev else 0)
        rc = self.L.eepacc_run(argv, resv, iw, None, self.mem)
        if rc:
            raise EepaccError(f"eepacc_run failed ({rc}): {self._err()}")
        return log, kpi

    def counters(self):
        out = np.zeros((self.B, 4))
        self.L.eepacc_mem_get_counters(self.mem, out.ctypes.data_as(_dp))
        return out

    def launch_count(self):
        return int(self.L.eepacc_mem_launch_count(self.mem))

    def enable_timing(self, on=True):
        self.L.eepacc_mem_enable_timing(self.mem, 1 if on else 0)

    def set_groups(self, groups):
        """Number of slices `run` pipelines the batch as (0 = automatic)."""
        if self.L.eepacc_mem_set_groups(self.mem, int(groups)):
            raise EepaccError(self._err())

    def groups(self):
        return int(self.L.eepacc_mem_get_groups(self.mem))

    def set_solver(self, solver="qpoases", **options):
        """Closed-loop QP solver (the reference's OPTsettings.solverToUse, MC/Settings.m:112): "qpoases" (default) =
        condensing + dense online active set (K3 / K3r); "ocp" = the same QP kept stage-structured and solved by a
        Riccati-based interior-point method, one vehicle per thread (K6; FBMPC / ABMPC).  options for "ocp":
        mu0, eps_mu, eps_p, eps_d, eps_accept, maxit (include/eepacc.h)."""
        code = {"qpoases": 0, "dense": 0, "ocp": 1, "hpipm": 1}[solver] if isinstance(solver, str) else int(solver)
        if self.L.eepacc_mem_set_solver(self.mem, code):
            raise EepaccError(self._err())
        if options:
            keys = ["mu0", "eps_mu", "eps_p", "eps_d", "eps_accept", "maxit"]
            bad = set(options) - set(keys)
            if bad:
                raise EepaccError(f"unknown solver options {sorted(bad)}")
            o = np.array([float(options.get(k, np.nan)) for k in keys])
            if self.L.eepacc_mem_set_ocp_options(self.mem, o.ctypes.data_as(_dp)):
                raise EepaccError(self._err())

    def solver(self):
        return {0: "qpoases", 1: "ocp"}.get(int(self.L.eepacc_mem_get_solver(self.mem)), "?")

    def set_fused(self, on=True):
        """run(): False (default) = per-step launches over pipelined slices; True = one fused closed-loop launch
        per cycle after the cold-start step."""
        if self.L.eepacc_mem_set_fused(self.mem, 1 if on else 0):
            raise EepaccError(self._err())

    def fused(self):
        return bool(self.L.eepacc_mem_get_fused(self.mem))

    def cycle_timing(self):
        """-> dict(ms, grid, ctas_per_sm) of the last fused closed-loop launch (CUDA events on its stream)."""
        out = np.zeros(7)
        if self.L.eepacc_mem_get_cycle_timing(self.mem, out.ctypes.data_as(_dp)):
            raise EepaccError(self._err())
        return dict(ms=float(out[0]), grid=int(out[1]), ctas_per_sm=int(out[2]),
                    cycles=dict(build=out[3], qp=out[4], handover_qp=out[5], post=out[6]))

    def timing(self, n_steps):
        """-> dict(build_ms, qp_ms, post_ms, steps, step_ms[n_steps * groups]) since the last reset/run start;
        the ms sums are kernel time per slice, which overlaps between slices when groups > 1."""
        tot = np.zeros(4)
        st = np.zeros(n_steps * max(self.groups(), 1))
        rc = self.L.eepacc_mem_get_timing(self.mem, tot.ctypes.data_as(_dp), st.ctypes.data_as(_dp), st.size)
        if rc:
            raise EepaccError(self._err())
        return dict(build_ms=tot[0], qp_ms=tot[1], post_ms=tot[2], steps=int(tot[3]), step_ms=st)

    def debug_trace(self, cap=4096):
        """Developer aid: first call enables tracing, later calls return the last solve's trace."""
        out = np.zeros((cap, 4))
        self.L.eepacc_mem_debug_trace(self.mem, out.ctypes.data_as(_dp), cap)
        return out[out[:, 2] > 0]

    def debug_phases(self):
        out = np.zeros(28)
        self.L.eepacc_mem_debug_phases(self.mem, out.ctypes.data_as(_dp))
        return out

    def debug_fetch(self, what, count):
        """developer aid: per-step device buffer -> host (0 sblk, 1 g, 2 lba, 3 uba, 4 x, 5 status, 6 iter, 7 cost)"""
        out = np.zeros(int(count))
        self.L.eepacc_mem_debug_fetch.restype = C.c_longlong
        got = self.L.eepacc_mem_debug_fetch(self.mem, int(what), out.ctypes.data_as(_dp), C.c_longlong(out.size))
        if got < 0:
            raise EepaccError(f"eepacc_mem_debug_fetch({what}) failed ({got}): {self._err()}")
        return out[:got]

    def working_set(self):
        """int8 [B, m + n]: working set after the last solve (-1 lower, +1 upper, 0 inactive; rows, then the
        implicit variable bounds) -- what the warm start of the next step begins from."""
        return self.debug_fetch(10, self.B * (self.m + self.n)).reshape(self.B, self.m + self.n).astype(np.int8)

    def kpi_device_ptr(self):
        return self.L.eepacc_mem_kpi_device_ptr(self.mem)


def _run_opt(ctrl, params, routes, lead, n_steps=None, device=-1):
    lead = np.asarray(lead, dtype=np.float64)
    if lead.ndim == 2:
        lead = lead[:, None, :]
    B = lead.shape[1]
    n_steps = n_steps or lead.shape[0]
    eng = Engine(ctrl, params, routes, B=B, device=device)
    try:
        eng.set_lead(lead)
        log, kpi = eng.run(n_steps)
        ctr = eng.counters()
    finally:
        eng.close()
    sol = dict(s_opt=log[:, :, 0], v_opt=log[:, :, 1], Fm_opt=log[:, :, 2], Fb_opt=log[:, :, 3],
               xi_v_opt=log[:, :, 4], xi_h_opt=log[:, :, 5], xi_s_opt=log[:, :, 6], xi_f_opt=log[:, :, 7],
               a_opt=log[:, :, 8], exitMessage=(log[:, :, 9] != 0), iter=log[:, :, 10], cost=log[:, :, 11],
               kpi=kpi, counters=ctr)
    Ts = params.Ts
    rpm = (30 / np.pi) * sol["v_opt"] * params.phi                 # RunOpt_FBMPC.m:334-339
    sol["rpm_opt"] = rpm
    sol["E_opt_end"] = kpi[:, 0]
    sol["j_opt"] = np.diff(sol["a_opt"], axis=0) / Ts
    return sol


def RunOpt_ABMPC(params, routes, lead, **kw):
    """RunOpt_ABMPC.m:1 for a batch: lead [n_steps, B, 2] (or [n_steps, 2])."""
    return _run_opt(S.AB, params, routes, lead, **kw)


def RunOpt_FBMPC(params, routes, lead, **kw):
    """RunOpt_FBMPC.m:1 for a batch."""
    return _run_opt(S.FB, params, routes, lead, **kw)


def RunOpt_BLMPC(params, routes, lead, **kw):
    """RunOpt_BLMPC.m:1 for a batch."""
    return _run_opt(S.BL, params, routes, lead, **kw)


def RunOpt_TVMPC(params, routes, n_steps, B=1, **kw):
    """RunOpt_TVMPC.m:1 for a batch (no lead vehicle)."""
    lead = np.zeros((n_steps, B, 2))
    return _run_opt(S.TV, params, routes, lead, **kw)
