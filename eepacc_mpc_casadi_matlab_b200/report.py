"""Result report of a batch of closed-loop runs: the numbers Main.m prints, one per vehicle.

Restates Main.m:132-262 (cut-off index, feasibility, distance, energy, speed-limit error, travel time and
comfort metrics at the cut-off distance) and RunOpt_FBMPC.m:334-339 (rpm, power, cumulative energy, jerk) for
arrays shaped [n_steps, B] -- the `sol` dict of api.RunOpt_*MPC.  This is tear-down code on the host (numpy over
the logged first-stage values); the per-vehicle KPI block the engine computes on the device (`sol["kpi"]`,
gathered over NCCL by sharding.gather_kpis) holds the same energy / travel-time / acceleration entries and
`tests/` check the two against each other.
"""
import numpy as np

PCT = (50, 90, 99)


def power5(b, Fm, rpm):
    """GetMotorPower_FifthOrderSurface.m:16-20: full bivariate quintic in (Fm, rpm), 21 coefficients."""
    x, y = np.asarray(Fm, dtype=np.float64), np.asarray(rpm, dtype=np.float64)
    out = np.zeros(np.broadcast(x, y).shape)
    k = 0
    for deg in range(6):
        for j in range(deg + 1):                      # x^deg, x^(deg-1) y, ..., y^deg
            out = out + b[k] * x ** (deg - j) * y ** j
            k += 1
    return out


def interp_pwa(x, doms, vals):
    """InterpPWA.m:14-27: clamp outside the domain, linear inside."""
    return np.interp(x, np.asarray(doms, dtype=np.float64), np.asarray(vals, dtype=np.float64))


def cutoff_index(s_opt, cut_off_dist):
    """Main.m:136-148: first MATLAB index i with s(i) < cutOffDist < s(i+1), else length-1.  s_opt [n, B] -> int[B]."""
    s = np.asarray(s_opt, dtype=np.float64)
    if s.ndim == 1:
        s = s[:, None]
    n, B = s.shape
    hit = (s[:-1] < cut_off_dist) & (s[1:] > cut_off_dist)
    first = np.argmax(hit, axis=0) + 1                # 1-based
    return np.where(hit.any(axis=0), first, n - 1).astype(np.int64)


def _rms(a, cnt):
    return np.sqrt(np.sum(a * a, axis=0) / cnt)


def kpi_report(sol, params, routes=None, lead=None):
    """Per-vehicle report.  sol: dict with s_opt, v_opt, Fm_opt, a_opt, exitMessage (+ xi_*_opt) as [n, B] arrays.
    routes: Route or sequence of B Routes (for the speed-limit error); lead [n, B, 2] (for headway figures).
    Returns a dict of float64[B] arrays; keys follow Main.m's print-out."""
    s = np.atleast_2d(np.asarray(sol["s_opt"], dtype=np.float64).T).T
    v = np.asarray(sol["v_opt"], dtype=np.float64).reshape(s.shape)
    Fm = np.asarray(sol["Fm_opt"], dtype=np.float64).reshape(s.shape)
    a = np.asarray(sol["a_opt"], dtype=np.float64).reshape(s.shape)
    n, B = s.shape
    Ts = params.Ts
    rpm = (30.0 / np.pi) * v * params.phi                                       # RunOpt_FBMPC.m:334
    P = power5(list(params.b_fifth), Fm, rpm)                                   # :335
    E = Ts * np.cumsum(P, axis=0)                                               # :336-338
    j = np.diff(a, axis=0) / Ts                                                 # :339
    di = cutoff_index(s, params.cutOffDist)                                     # Main.m:136-148 (1-based)
    cols = np.arange(B)
    rows = np.arange(n)[:, None]
    upto = rows < di[None, :]                                                   # MATLAB 1:distInd
    upto_j = upto[:-1] if n > 1 else upto
    at_cut = np.clip(di - 2, 0, n - 1)                                          # MATLAB index distInd-1
    rep = {
        "bad_exit_messages": np.sum(np.asarray(sol["exitMessage"]).reshape(s.shape) != 0, axis=0).astype(np.float64),
        "distance_km": s[-1] / 1e3,                                             # :200-204
        "energy_kWh": E[-1] / 3.6e6,                                            # :206-210
        "cutoff_index": di.astype(np.float64),
        "energy_at_cutoff_kWh": E[at_cut, cols] / 3.6e6,                        # :228-233
        "travel_time_at_cutoff_s": 0.1 * np.floor(di * Ts * 10 + 0.5),          # :221-225,235-239
        "a_max": np.max(np.where(upto, a, -np.inf), axis=0),                    # :241-245
        "a_min": np.min(np.where(upto, a, np.inf), axis=0),
        "a_rms": _rms(np.where(upto, a, 0.0), di),
        "j_max": np.max(np.where(upto_j, j, -np.inf), axis=0),                  # :247-250 (j has n-1 entries)
        "j_min": np.min(np.where(upto_j, j, np.inf), axis=0),
        "j_rms": _rms(np.where(upto_j, j, 0.0), np.minimum(di, n - 1)),
        "v_max": v.max(axis=0),
    }
    if routes is not None:
        rl = [routes] * B if hasattr(routes, "n_sl") else list(routes)
        vlim = np.array([interp_pwa(params.cutOffDist, r.s_speedLim[:r.n_sl], r.v_speedLim[:r.n_sl]) for r in rl])
        rep["speed_limit_error_at_cutoff"] = vlim - v[at_cut, cols]             # Main.m:133,212-217
    for name in ("xi_v_opt", "xi_h_opt", "xi_s_opt", "xi_f_opt"):
        if name in sol:                                                         # constraint-violation counts
            rep["steps_" + name[:-4] + "_active"] = np.sum(
                np.asarray(sol[name]).reshape(s.shape) > 1e-6, axis=0).astype(np.float64)
    if lead is not None:
        ld = np.asarray(lead, dtype=np.float64).reshape(n, B, 2)
        gap = ld[:, :, 0] - s
        fin = np.isfinite(gap)
        rep["min_gap_m"] = np.min(np.where(fin, gap, np.inf), axis=0)
        rep["steps_below_h_min"] = np.sum(fin & (gap < params.h_min), axis=0).astype(np.float64)
        rep["steps_below_tau_min"] = np.sum(fin & (gap < params.tau_min * v), axis=0).astype(np.float64)
    rep["_series"] = dict(rpm_opt=rpm, P_opt=P, E_opt=E, j_opt=j)
    return rep


def batch_summary(rep, pct=PCT):
    """Mean, min, max and percentiles over the batch for every scalar entry of kpi_report()."""
    out = {}
    for k, val in rep.items():
        if k.startswith("_"):
            continue
        x = np.asarray(val, dtype=np.float64)
        x = x[np.isfinite(x)]
        if x.size == 0:
            continue
        out[k] = dict(mean=float(x.mean()), min=float(x.min()), max=float(x.max()),
                      **{f"p{q}": float(np.percentile(x, q)) for q in pct})
    return out


def format_report(reports, use_case_num=0, cut_off_dist=None, vehicle=0):
    """Main.m:188-251 as text for one vehicle of each controller.  reports: {"Force-based MPC": rep, ...}."""
    L = [f"=== RESULTS of UC{use_case_num} ===", "", "MPC solver: dense active set (one vehicle per CTA)", "", "Feasibility:"]
    for nm, r in reports.items():
        L.append(f"   {nm}: {int(r['bad_exit_messages'][vehicle])} bad exit messages")
    blocks = (("Distance traveled:", "distance_km", "km"), ("Energy consumption:", "energy_kWh", "kWh"),
              ("Error relative to the speed limit at cut-off:", "speed_limit_error_at_cutoff", "m/s"),
              ("Energy consumption at cut-off:", "energy_at_cutoff_kWh", "kWh"),
              ("Travel time at cut-off:", "travel_time_at_cutoff_s", "s"))
    for title, key, unit in blocks:
        if not all(key in r for r in reports.values()):
            continue
        L += ["", title if cut_off_dist is None else title.replace("cut-off", f"{cut_off_dist / 1e3:g} km")]
        for nm, r in reports.items():
            L.append(f"   {nm}: {r[key][vehicle]:.5g} {unit}")
    L += ["", "Comfort metrics at cut-off:"]
    for q, unit in (("a", "m/s2"), ("j", "m/s3")):
        for nm, r in reports.items():
            L.append(f"   {nm}: max. {q} = {r[q + '_max'][vehicle]:.5g}{unit}, min. {q} = {r[q + '_min'][vehicle]:.5g}{unit}, "
                     f"rms. {q} = {r[q + '_rms'][vehicle]:.5g}{unit}")
    return "\n".join(L)
