/* nlp_oracle.c -- TEST INFRASTRUCTURE: a plain C restatement of the NLP that MATLAB_CasADi/RunOpt_NLP.m hands to
 * nlpsol('ipopt') -- objective J(z), constraint vector g(z) and the bounds lbz/ubz/lbg/ubg -- in the reference's own
 * variable and row order (multiple shooting, RunOpt_NLP.m:296-497).  Only tests/, __graft_entry__.smoke() and
 * bench.py's CPU arm may use it, and only as the checker.
 *
 * What it is for: the solver itself (IPOPT inside CasADi's binary-only plugin) is absent from /root/reference, so the
 * product's NLP path (csrc/k7_nlp.cuh) is checked by (i) evaluating THIS restatement at the reference's own saved
 * optimum (savedNLPsol.mat -> tests/golden/nlp_*.npz: all rows within ipopt's 1e-8 bound relaxation) -- that pins the
 * restatement; (ii) evaluating it at the product's answer: feasibility, objective against the saved optimum's.
 * It shares no code with csrc/k7_nlp.cuh (different variables: the kernel eliminates theta and carries a force state).
 */
#include <math.h>
#include <string.h>
#include "eepacc_oracle.h"

/* casadi.interpolant('LUT','linear',{xs},ys) (RunOpt_NLP.m:66-80,112,155,184): linear inside, linear extrapolation outside */
double orc_lut_linear(double x, const double* xs, const double* ys, int n)
{
    if (n < 2) return n == 1 ? ys[0] : 0.0;
    int i = 0;
    while (i < n - 2 && x >= xs[i + 1]) ++i;
    return ys[i] + (ys[i + 1] - ys[i]) / (xs[i + 1] - xs[i]) * (x - xs[i]);
}

int orc_nlp_rows_per_stage(const eepacc_route* r) { return 4 + 17 + 2 * r->n_tl; }

/* z: [X0 | U0 X1 | ... ] with X = (s, v, theta, j), U = (Fm, Fb, xi_v, xi_h, xi_s, xi_f); s_tv: [N + 1].
 * Outputs: *f; g, lbg, ubg: [N * rows_per_stage]; lbz, ubz: [4 + 10 N] (any output may be NULL).  Returns the row count. */
int orc_nlp_eval(const eepacc_params* p, const eepacc_route* r, const double* s_velInc, const double* v_velInc, int n_vi,
                 const double* s_stop, const double* v_stop, int n_stop, const double* s_tv, int N, const double* z,
                 double* f, double* g, double* lbg, double* ubg, double* lbz, double* ubz)
{
    const double m = p->m * r->mass_scale, Ts = p->Ts, INF = INFINITY;
    const double Fm_min = -p->phi * p->T_m_max / p->eta_TF, Fm_max = p->phi * p->T_m_max * p->eta_TF;   /* :195-196 */
    const double gv[4] = { 0, 5, 20, 25 }, amin[4] = { -4, -4, -2, -2 }, amax[4] = { 5, 5, 3.5, 3.5 },
                 jmag[4] = { 5, 5, 2.5, 2.5 };                                                             /* :74-80 */
    double slope_sum = 0.0;
    for (int i = 0; i < r->n_slope; ++i) slope_sum += r->slope[i];
    const int rows = orc_nlp_rows_per_stage(r);
    double J = 0.0;
    int ir = 0;
#define ROW(lo, val, hi) do { if (g) g[ir] = (val); if (lbg) lbg[ir] = (lo); if (ubg) ubg[ir] = (hi); ++ir; } while (0)
    if (lbz) {                                                                                            /* :308-312 */
        const double th0 = orc_lut_linear(p->s_init, r->s_slope, r->slope, r->n_slope);
        const double xi[4] = { p->s_init, p->v_init, th0, 0.0 };
        for (int i = 0; i < 4; ++i) { lbz[i] = xi[i]; if (ubz) ubz[i] = xi[i]; }
    }
    double Uprev[6] = { 0, 0, 0, 0, 0, 0 };                                                               /* :338 */
    const double* Xprev = z;
    for (int k = 0; k < N; ++k) {
        const double* U = z + 4 + 10 * k;
        const double* X = U + 6;
        if (lbz) {                                                                                        /* :346-348, :364-366 */
            double* lo = lbz + 4 + 10 * k; double* hi = ubz + 4 + 10 * k;
            lo[0] = Fm_min; hi[0] = Fm_max; lo[1] = -INF; hi[1] = 0.0;
            for (int i = 2; i < 6; ++i) { lo[i] = 0.0; hi[i] = INF; }
            lo[6] = 0.0; hi[6] = p->s_goal; lo[7] = 0.0; hi[7] = p->v_max; lo[8] = -INF; hi[8] = INF; lo[9] = -INF; hi[9] = INF;
        }
        double xf[4], qf;
        orc_F(p, r->mass_scale, Xprev, U, xf, &qf);                                                       /* :351-352 */
        J += qf;
        const double s = X[0], v = X[1], th = X[2], j = X[3];
        ROW(0, xf[0] - s, 0);                                                                             /* :370-372 */
        ROW(0, xf[1] - v, 0);
        if (slope_sum < 1e-1) ROW(0, th, 0);                                                              /* :377-382 */
        else ROW(0, th - orc_lut_linear(s, r->s_slope, r->slope, r->n_slope), 0);
        const double cth = cos(th), sth = sin(th);
        const double al = U[0] + U[1] - p->zeta_a * v * v - p->c_r * m * p->g * cth - m * p->g * sth;
        const double alp = Uprev[0] + Uprev[1] - p->zeta_a * Xprev[1] * Xprev[1] - p->c_r * m * p->g * cos(Xprev[2]) -
                           m * p->g * sin(Xprev[2]);
        ROW(0, j - (al - alp) / (p->lambda * m * Ts), 0);                                                 /* :384-387 */
        const double Fm = U[0], Fb = U[1], xiv = U[2], xih = U[3], xis = U[4], xif = U[5];
        ROW(0, Fm * v + p->P_m_max / p->eta_TF + xif, INF);                                               /* :390-392 */
        ROW(-INF, Fm * v - p->P_m_max * p->eta_TF - xif, 0);                                              /* :394-396 */
        ROW(0, Fm + Fb + p->mu * m * p->g * cth + xif, INF);                                              /* :399-401 */
        ROW(-INF, Fm + Fb - p->mu * m * p->g * cth - xif, 0);                                             /* :403-405 */
        const double rear = p->h_g * p->lambda * al / (p->lambda * m) + p->h_g * p->zeta_a / m * v * v +
                            p->g * (p->L_f * cth + p->h_g * sth);
        ROW(0, p->L / (p->mu * m) * Fm + rear + xif, INF);                                                /* :408-412 */
        ROW(-INF, p->L / (p->mu * m) * Fm - rear - xif, 0);                                               /* :414-418 */
        const double a = al / (p->lambda * m);
        ROW(0, a - orc_lut_linear(v, gv, amin, 4) + xif, INF);                                            /* :421-424 */
        ROW(-INF, a - orc_lut_linear(v, gv, amax, 4) - xif, 0);                                           /* :426-429 */
        ROW(0, j + orc_lut_linear(v, gv, jmag, 4) + xif, INF);                                            /* :432-434 */
        ROW(-INF, j - orc_lut_linear(v, gv, jmag, 4) - xif, 0);                                           /* :436-438 */
        ROW(-INF, v - orc_lut_linear(s, r->s_speedLim, r->v_speedLim, r->n_sl) - xif, 0);                 /* :441-443 */
        const double kap = fabs(orc_lut_linear(s, r->s_curv, r->curvature, r->n_curv));
        ROW(-INF, v - (kap > 1e-15 ? p->alpha_TTL * pow(kap, -1.0 / 3.0) : 1e5) - xif, 0);                /* :446-448 */
        ROW(-INF, v - orc_lut_linear(s, s_stop, v_stop, n_stop) - xis, 0);                                /* :451-453 */
        for (int i = 0; i < r->n_tl; ++i) {                                                               /* :456-470 */
            const double* tl = r->TLLoc[i];
            const double D = p->stopRefDist, inc = D * p->stopRefVelSlope;
            const double xs[3] = { tl[0] - D, tl[0], tl[0] + D }, ys[3] = { inc, p->TLstopVel, inc };
            const double vp = orc_lut_linear(s, xs, ys, 3);
            const double per = tl[2] + tl[3], tt = k * Ts - tl[1];
            const double st = (tt - floor(tt / per) * per) < tl[2] ? .2 : 1e3;                            /* :147-151 */
            ROW(-INF, v - vp - st - xis, 0);
            ROW(0, v + vp + 1e3 - 10 - st + xis, INF);
        }
        ROW(0, v - orc_lut_linear(s, s_velInc, v_velInc, n_vi) + xiv, INF);                               /* :473-475 */
        ROW(-INF, s, s_tv[k] - p->h_min);                                                                 /* :478-480, s_tv(k+1) 1-based */
        ROW(-INF, s + p->tau_min * v - xis, s_tv[k]);
        const double A_hwp = 2, T_hwp = 2, G_hwp = -0.0246 * T_hwp + 0.010819;                            /* :483-488 */
        ROW(-INF, s + v * T_hwp + v * v * G_hwp - xih, s_tv[k] - A_hwp);
        memcpy(Uprev, U, sizeof Uprev);
        Xprev = X;
    }
#undef ROW
    if (f) *f = J;
    (void)rows;
    return ir;
}
