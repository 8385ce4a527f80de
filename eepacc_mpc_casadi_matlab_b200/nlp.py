"""Host side of the batched NLP benchmark: `RunOpt_NLP.m` for many scenarios at once.

Mirrors the reference's call (MATLAB_CasADi/RunOpt_NLP.m:499-510)

    solver = nlpsol('solver', 'ipopt', prob, opts);  sol = solver('x0', z0, 'lbx', ..., 'ubg', ...);

as ``NlpSolver(params, routes, n_steps).solve(s_tv, x0=None) -> NlpSolution`` over the C-ABI of
``include/eepacc_nlp.h`` (kernel ``csrc/k7_nlp.cu``).  The lookup tables that RunOpt_NLP.m builds before it states
the problem (:86-184: stop velocity profile, velocity-incentive PWA) are set-up work on the host (`nlp_tables`).
There is no CPU solver here: without a CUDA device `NlpSolver` raises.
"""
import ctypes as C

import numpy as np

from . import api
from .settings import Params, Route
from .usecase import simplify_pwa

MAX_VI, MAX_STOPV, N_STATS = 48, 24, 8        # eepacc_nlp.h
STATUS = {0: "Solve_Succeeded", 1: "Maximum_Iterations_Exceeded", 2: "Line_Search_Failed", 3: "Numerical_Error",
          4: "Solved_To_Acceptable_Level"}


class NlpTables(C.Structure):
    _fields_ = [("s_velInc", C.c_double * MAX_VI), ("v_velInc", C.c_double * MAX_VI),
                ("s_stop", C.c_double * MAX_STOPV), ("v_stop", C.c_double * MAX_STOPV),
                ("n_vi", C.c_int), ("n_stop", C.c_int)]


class NlpOpts(C.Structure):
    _fields_ = [("tol", C.c_double), ("mu0", C.c_double), ("obj_scale", C.c_double), ("mu_min", C.c_double),
                ("acceptable_kkt", C.c_double), ("max_iter", C.c_int), ("pad_", C.c_int)]


# ---------------------------------------------------------------------------------------------------------------
# set-up tables
# ---------------------------------------------------------------------------------------------------------------
def _interp_clamped(x, xs, ys):
    return float(np.interp(x, xs, ys))          # InterpPWA.m: clamps outside the domain


def pwa_min(a_dom, a_val, b_dom, b_val):
    """Point-wise minimum of two PWA functions (what minPWA.m computes): all break points of both plus the crossing
    points of every pair of overlapping segments, each valued with the smaller function."""
    lo, hi = min(a_dom[0], b_dom[0]), max(a_dom[-1], b_dom[-1])
    a_dom, a_val = [lo] + list(a_dom) + [hi], [a_val[0]] + list(a_val) + [a_val[-1]]
    b_dom, b_val = [lo] + list(b_dom) + [hi], [b_val[0]] + list(b_val) + [b_val[-1]]
    pts = sorted(set(a_dom) | set(b_dom))
    out = set(pts)
    for x0, x1 in zip(pts[:-1], pts[1:]):
        d0 = _interp_clamped(x0, a_dom, a_val) - _interp_clamped(x0, b_dom, b_val)
        d1 = _interp_clamped(x1, a_dom, a_val) - _interp_clamped(x1, b_dom, b_val)
        if d0 * d1 < 0:
            out.add(x0 + (x1 - x0) * d0 / (d0 - d1))
    dom = sorted(out)
    val = [min(_interp_clamped(x, a_dom, a_val), _interp_clamped(x, b_dom, b_val)) for x in dom]
    return dom, val


def saturate_slope(dom, val, c):
    """SaturateSlopePWA.m: a segment steeper than c is flattened to slope c by moving its up-hill break point
    outwards (a rise ends later, a drop starts earlier); break points overtaken by the move are dropped."""
    dom, val = list(map(float, dom)), list(map(float, val))
    for _ in range(2):
        for i in range(1, len(dom)):
            dx = dom[i] - dom[i - 1]
            if dx <= 0:
                continue
            sl = (val[i] - val[i - 1]) / dx
            if sl > c:
                dom[i] = dom[i - 1] + (val[i] - val[i - 1]) / c
            elif sl < -c:
                dom[i - 1] = dom[i] + (val[i] - val[i - 1]) / c
        keep_d, keep_v = [dom[0]], [val[0]]
        for d, v in zip(dom[1:], val[1:]):
            while len(keep_d) > 1 and d < keep_d[-1]:
                keep_d.pop(); keep_v.pop()
            if d < keep_d[-1]:
                keep_d[-1] = d
            keep_d.append(d); keep_v.append(v)
        dom, val = keep_d, keep_v
    return dom, val


def velocity_incentive(route, alpha_TTL):
    """RunOpt_NLP.m:160-183: min(speed limit, two-thirds-law curve speed), slopes saturated at 0.5 (m/s)/m, duplicates
    removed, SimplifyPWA."""
    n_sl, n_cv = route.n_sl, route.n_curv
    s_sl, v_sl = list(route.s_speedLim[:n_sl]), list(route.v_speedLim[:n_sl])
    s_cv = list(route.s_curv[:n_cv])
    with np.errstate(divide="ignore"):
        v_cv = [min(alpha_TTL * abs(k) ** (-1.0 / 3.0), 1e5) if k != 0 else 1e5 for k in route.curvature[:n_cv]]
    d, v = pwa_min(s_sl, v_sl, s_cv, v_cv)
    d, v = simplify_pwa(d, v)
    d, v = saturate_slope(d, v, 0.5)
    dd, vv = [d[0]], [v[0]]
    for a, b in zip(d[1:], v[1:]):
        if a != dd[-1]:
            dd.append(a); vv.append(b)
    d, v = simplify_pwa(dd, vv)
    return d, v


def stop_profile(route, params):
    """RunOpt_NLP.m:86-112: 3 points per stop, the merge of close stops, [0, 1] -> 1e5 without stops."""
    stops = sorted(route.stopLoc[:route.n_stop])
    D = params.stopRefDist
    inc = D * params.stopRefVelSlope
    s, v = [], []
    for loc in stops:
        s += [loc - D, loc, loc + D]
        v += [inc, params.stopVel, inc]
    for i in range(len(v) - 1):
        if s[i + 1] <= s[i]:
            gap = s[i] - s[i + 1]
            mid = .5 * gap + s[i + 1]
            v[i] = v[i + 1] = inc / (1 + D / gap) if gap > 0 else 0.0
            s[i], s[i + 1] = mid - 1, mid + 1
    if not stops:
        s, v = [0.0, 1.0], [1e5, 1e5]
    return s, v


def nlp_tables(route, params):
    t = NlpTables()
    d, v = velocity_incentive(route, params.alpha_TTL)
    sd, sv = stop_profile(route, params)
    if len(d) > MAX_VI or len(sd) > MAX_STOPV:
        raise api.EepaccError("a lookup table of the NLP needs more points than eepacc_nlp_tables holds")
    t.n_vi, t.n_stop = len(d), len(sd)
    t.s_velInc[:len(d)] = d
    t.v_velInc[:len(d)] = v
    t.s_stop[:len(sd)] = sd
    t.v_stop[:len(sd)] = sv
    return t


# ---------------------------------------------------------------------------------------------------------------
# the solver object
# ---------------------------------------------------------------------------------------------------------------
class NlpSolution:
    """Fields of RunOpt_NLP.m's optSol (:533-556), each [B][...]."""

    def __init__(self, z, f, stats, n_steps, ms):
        B = z.shape[0]
        N = n_steps
        self.z, self.f, self.stats, self.ms = z, f, stats, ms
        X = np.concatenate([z[:, :4].reshape(B, 1, 4), z[:, 4:].reshape(B, N, 10)[:, :, 6:]], axis=1)
        U = z[:, 4:].reshape(B, N, 10)[:, :, :6]
        self.s_opt, self.v_opt, self.theta_opt, self.j_opt = (X[:, :, i] for i in range(4))
        self.Fm_opt, self.Fb_opt, self.xi_v_opt, self.xi_h_opt, self.xi_s_opt, self.xi_f_opt = (U[:, :, i] for i in range(6))
        self.status = stats[:, 0].astype(int)
        self.iterations = stats[:, 1].astype(int)
        self.kkt_error = stats[:, 2]
        self.exitMessage = [STATUS.get(int(s), "?") for s in self.status]


class NlpSolver:
    def __init__(self, params, routes, n_steps, B=None, device=0, tables=None):
        self.L = api.load_library()
        L = self.L
        L.eepacc_nlp_last_error.restype = C.c_char_p
        L.eepacc_nlp_last_ms.restype = C.c_double
        L.eepacc_nlp_nz.restype = C.c_longlong
        L.eepacc_nlp_workspace_bytes.restype = C.c_longlong
        L.eepacc_nlp_workspace_bytes.argtypes = [C.c_longlong, C.c_int]
        L.eepacc_nlp_setup.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int]
        L.eepacc_nlp.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        assert L.eepacc_nlp_sizeof_tables() == C.sizeof(NlpTables)
        if isinstance(routes, Route):
            routes = [routes] * (B or 1)
        self.B, self.N = len(routes), int(n_steps)
        self.params = params
        R = (Route * self.B)(*routes)
        tabs = tables if tables is not None else [nlp_tables(r, params) for r in routes]
        T = (NlpTables * self.B)(*tabs)
        self.mem = L.eepacc_nlp_checkout()
        if self.mem < 0:
            raise api.EepaccError(L.eepacc_nlp_last_error().decode())
        rc = L.eepacc_nlp_setup(self.mem, C.byref(params), R, T, self.B, self.N, device)
        if rc:
            msg = L.eepacc_nlp_last_error().decode()
            L.eepacc_nlp_release(self.mem)
            self.mem = -1
            raise api.EepaccError(f"eepacc_nlp_setup failed ({rc}): {msg}")

    def set_options(self, **kw):
        o = NlpOpts()
        self.L.eepacc_nlp_default_opts(C.byref(o))
        for k, v in kw.items():
            setattr(o, k, v)
        if self.L.eepacc_nlp_set_opts(self.mem, C.byref(o)):
            raise api.EepaccError(self.L.eepacc_nlp_last_error().decode())

    def workspace_bytes(self):
        return int(self.L.eepacc_nlp_workspace_bytes(self.B, self.N))

    def solve(self, s_tv, x0=None):
        """s_tv: [B][n_steps + 1] host array (OPTsettings.s_tv); x0: optional [B][4 + 10 n_steps] guess."""
        B, N = self.B, self.N
        nz = 4 + 10 * N
        p = np.ascontiguousarray(s_tv, dtype=np.float64)
        if p.shape != (B, N + 1):
            raise api.EepaccError(f"s_tv must be [{B}][{N + 1}]")
        z = np.zeros((B, nz)); f = np.zeros(B); st = np.zeros((B, N_STATS))
        g = None
        if x0 is not None:
            g = np.ascontiguousarray(x0, dtype=np.float64)
            if g.shape != (B, nz):
                raise api.EepaccError(f"x0 must be [{B}][{nz}]")
        arg = (C.c_void_p * 2)(g.ctypes.data if g is not None else None, p.ctypes.data)
        res = (C.c_void_p * 3)(z.ctypes.data, f.ctypes.data, st.ctypes.data)
        iw = (C.c_longlong * 2)(B, 0)
        rc = self.L.eepacc_nlp(arg, res, iw, None, self.mem)
        if rc:
            raise api.EepaccError(f"eepacc_nlp failed ({rc}): {self.L.eepacc_nlp_last_error().decode()}")
        return NlpSolution(z, f, st, N, float(self.L.eepacc_nlp_last_ms(self.mem)))

    def close(self):
        if getattr(self, "mem", -1) >= 0:
            self.L.eepacc_nlp_release(self.mem)
            self.mem = -1

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def RunOpt_NLP(params, routes, s_tv, n_steps=None, device=0, **options):
    """The reference's entry point, batched: returns an NlpSolution (optSol of RunOpt_NLP.m:533-575)."""
    s_tv = np.atleast_2d(np.asarray(s_tv, dtype=np.float64))
    n_steps = n_steps or s_tv.shape[1] - 1
    if isinstance(routes, Route):
        routes = [routes] * s_tv.shape[0]
    S = NlpSolver(params, routes, n_steps, device=device)
    if options:
        S.set_options(**options)
    try:
        return S.solve(s_tv[:, :n_steps + 1])
    finally:
        S.close()


def summary_rows(sol):
    """[B][2 + N_STATS] rows a rank contributes to the gathered report: objective, travelled distance, then the stats columns."""
    return np.concatenate([sol.f[:, None], sol.s_opt[:, -1:], sol.stats], axis=1)


def RunOpt_NLP_sharded(params, routes, s_tv, rank, world, device=None, solve=None, **options):
    """Multi-GPU form: the scenarios are independent, so rank g of G solves the contiguous block
    sharding.shard_range(B, g, G) on its own GPU with no data-path collective; the only exchange is one all_gather of the
    per-scenario summary rows (`summary_rows`) over torch.distributed (NCCL on the box, gloo in the CPU test).
    Returns (this rank's NlpSolution, gathered rows [B][2 + N_STATS] in global scenario order).
    `solve(params, routes_block, s_tv_block)` replaces the GPU call in the CPU test of the host-side logic."""
    import torch
    from . import sharding
    s_tv = np.atleast_2d(np.asarray(s_tv, dtype=np.float64))
    B = s_tv.shape[0]
    if isinstance(routes, Route):
        routes = [routes] * B
    lo, hi = sharding.shard_range(B, rank, world)
    if solve is None:
        sol = RunOpt_NLP(params, routes[lo:hi], s_tv[lo:hi], device=rank if device is None else device, **options)
    else:
        sol = solve(params, routes[lo:hi], s_tv[lo:hi])
    rows = torch.from_numpy(summary_rows(sol))
    if torch.cuda.is_available() and solve is None:
        rows = rows.cuda(rank if device is None else device)
    return sol, sharding.gather_kpis(rows).cpu().numpy()
