/*
 * qp_oracle.c -- CPU ORACLE (test infrastructure, not a product path; see eepacc_oracle.h).
 *
 * Dense, warm-started, parametric (homotopy) primal-dual active-set QP solver with a
 * null-space (TQ + projected Cholesky) factorisation.  It stands where the reference calls
 *     sol = QPsolver('h',H,'g',c,'a',G,'lbx',..,'ubx',..,'lba',g_lb,'uba',g_ub)
 * (MC/RunOpt_ABMPC.m:252, RunOpt_FBMPC.m:278, RunOpt_BLMPC.m:230, RunOpt_TVMPC.m:208), i.e. for
 * CasADi 3.6.3's conic plugin 'qpoases' = qpOASES 3.2 (Ferreau, Kirches, Potschka, Bock, Diehl:
 * "qpOASES: a parametric active-set algorithm for quadratic programming", Math. Prog. Comp. 2014).
 * qpOASES' source is NOT in /root/reference (binary DLL only), so this file restates the
 * PUBLISHED algorithm, not its code:
 *   - online active-set strategy: homotopy from the previous QP's solution (x, y, working set)
 *     to the new data, step length from primal/dual ratio tests, one working-set change per stop;
 *   - null-space factorisation  A_W Q = [0 T],  Z'HZ = R'R  updated by Givens rotations;
 *   - all variables carry implicit "far" bounds (+-1e6, qpOASES enableFarBounds) and the cold
 *     start fixes every variable (qpOASES initialStatusBounds), so the projected Hessian is
 *     positive definite along the whole path even for semidefinite / indefinite / zero H;
 *   - a removal that would make Z'HZ lose definiteness is replaced by an exchange along the
 *     direction of non-positive curvature (plays the role of qpOASES' flipping bounds);
 *   - a linearly dependent addition is resolved by a dual ratio test (qpOASES ensureLI).
 * Tie-breaks are index ordered and deterministic; the CUDA solver follows the same rules.
 */
#include "eepacc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#define FARB      1.0e6      /* implicit variable bounds                                  */
#define RELAX     1.0e4      /* auxiliary-QP bound relaxation (qpOASES boundRelaxation)   */
#define BIG       1.0e19     /* |bound| >= BIG  <=> infinite                              */
#define EPS_LI    1.0e-10    /* linear-independence test, relative to |a|                 */
#define EPS_PD    1.0e-9     /* projected-Hessian pivot test, relative to max|diag H|     */
#define EPS_RATE  1.0e-12
#define EPS_ROW   1.0e-12    /* |a'dx| below EPS_ROW*|a|_1*|dx|_inf is numerically zero      */
#define FEAS_TOL  1.0e-10    /* a bound violated by less than this at the full step is not blocking (cf. qpOASES boundTolerance) */
#define LP_REG    1.0e-6
#define FLAT_TOL  1.0e-9     /* |g'd| <= FLAT_TOL |g| |d|: the gradient is orthogonal to d within round-off */
#define TIE_TOL   1.0e-11    /* ratio tests: candidates whose numerator lies within TIE_TOL*(scale) of the
                              * smallest ratio are TIED; the lowest constraint index among them is taken
                              * (two-pass Harris ratio test).  Makes the active-set path independent of
                              * round-off (summation order, FMA contraction) at degenerate vertices. */

/* ---- two-pass ratio test ----------------------------------------------------------------
 * candidate i blocks at t_i = num_i/den_i (num >= 0, den > 0); del_i = noise level of num_i.
 * pass 1: tmax = min_i (num_i + del_i)/den_i;  pass 2: among t_i <= tmax the lowest key wins.
 * With del = 0 this is the textbook "first minimum"; with del > 0 candidates that differ by
 * round-off only are resolved by index, identically in every arithmetic. */
typedef struct { double num, den, del; int key, aux; } rt_cand;
static int rt_select(const rt_cand* c, int nc)
{
    double tmax = INFINITY;
    int best = -1;
    for (int i = 0; i < nc; ++i) {
        const double inv = 1.0 / c[i].den;              /* one division per candidate, as in the kernels */
        double t = (c[i].num + c[i].del) * inv;
        if (t < tmax) tmax = t;
    }
    for (int i = 0; i < nc; ++i) {
        const double inv = 1.0 / c[i].den;
        if (c[i].num * inv <= tmax && (best < 0 || c[i].key < c[best].key)) best = i;
    }
    return best;
}

struct orc_qpws {
    int n, m, mt;
    int warm;
    int nA;
    int* widx;        /* [n]  constraint index per working-set position */
    int* st;          /* [mt] -1 lower, 0 inactive, +1 upper            */
    char* tabu;       /* [mt] anti-cycling: added during the current run of zero-length steps */
    double *x, *y, *Ax;
    double *g_cur, *lb_cur, *ub_cur, *lb_new, *ub_new;
    double *Q, *T, *R;                 /* n x n each */
    double *dx, *dyW, *Adx, *tmp1, *tmp2, *tmp3, *tmp4, *dg, *dlb, *dub;
    double *anorm;    /* [mt] L1 norm of every constraint row */
    const double *H, *A, *g_new;
    double hscale;
    double reg;       /* Tikhonov term for an all-zero Hessian (LP), see orc_qpsolve */
    double tie;       /* TIE_TOL, or 0 for the eps-regularised LPs (see orc_qpsolve) */
    double ynoise;    /* scale of the round-off carried by the multipliers: max|y| at set-up + sum of tau*max|dy| */
    int n_exch, n_refac;
    rt_cand* cand;    /* [2*mt + n] ratio-test candidates */
};

static double* dalloc(size_t n) { return (double*)calloc(n ? n : 1, sizeof(double)); }

orc_qpws* orc_qpws_new(int n, int m)
{
    orc_qpws* w = (orc_qpws*)calloc(1, sizeof *w);
    if (!w) return 0;
    w->n = n; w->m = m; w->mt = m + n;
    w->widx = (int*)calloc((size_t)n, sizeof(int));
    w->st = (int*)calloc((size_t)w->mt, sizeof(int));
    w->tabu = (char*)calloc((size_t)w->mt, 1);
    w->x = dalloc(n); w->y = dalloc(w->mt); w->Ax = dalloc(w->mt);
    w->g_cur = dalloc(n); w->lb_cur = dalloc(w->mt); w->ub_cur = dalloc(w->mt);
    w->lb_new = dalloc(w->mt); w->ub_new = dalloc(w->mt);
    w->Q = dalloc((size_t)n * n); w->T = dalloc((size_t)n * n); w->R = dalloc((size_t)n * n);
    w->dx = dalloc(n); w->dyW = dalloc(n); w->Adx = dalloc(w->mt);
    w->tmp1 = dalloc(w->mt); w->tmp2 = dalloc(w->mt); w->tmp3 = dalloc(w->mt); w->tmp4 = dalloc(w->mt);
    w->dg = dalloc(n); w->dlb = dalloc(w->mt); w->dub = dalloc(w->mt);
    w->anorm = dalloc(w->mt);
    w->cand = (rt_cand*)calloc((size_t)(2 * w->mt + n + 1), sizeof(rt_cand));
    return w;
}

void orc_qpws_free(orc_qpws* w)
{
    if (!w) return;
    free(w->widx); free(w->st); free(w->tabu); free(w->x); free(w->y); free(w->Ax); free(w->g_cur);
    free(w->lb_cur); free(w->ub_cur); free(w->lb_new); free(w->ub_new); free(w->Q); free(w->T);
    free(w->R); free(w->dx); free(w->dyW); free(w->Adx); free(w->tmp1); free(w->tmp2);
    free(w->tmp3); free(w->tmp4); free(w->dg); free(w->dlb); free(w->dub); free(w->anorm); free(w->cand);
    free(w);
}

void orc_qpws_reset(orc_qpws* w) { w->warm = 0; }

void orc_qpws_get_ws(const orc_qpws* w, int* ws) { memcpy(ws, w->st, sizeof(int) * (size_t)w->mt); }

/* ---- small dense helpers (column-major, leading dimension n) ---- */
#define Q_(i, j) Qm[(size_t)(j) * n + (i)]
#define R_(i, j) Rm[(size_t)(j) * n + (i)]
#define T_(i, c) Tm[(size_t)(i) * n + (c)]

static void get_row(const orc_qpws* w, int c, double* a)
{
    const int n = w->n, m = w->m;
    if (c < m) { for (int j = 0; j < n; ++j) a[j] = w->A[(size_t)j * m + c]; }
    else { memset(a, 0, sizeof(double) * (size_t)n); a[c - m] = 1.0; }
}

static void Hmul(const orc_qpws* w, const double* v, double* out)
{
    const int n = w->n;
    memset(out, 0, sizeof(double) * (size_t)n);
    for (int j = 0; j < n; ++j) {
        double vj = v[j];
        if (vj == 0.0) continue;
        const double* hc = w->H + (size_t)j * n;
        for (int i = 0; i < n; ++i) out[i] += hc[i] * vj;
    }
    if (w->reg != 0.0) for (int i = 0; i < n; ++i) out[i] += w->reg * v[i];
}

static void Amul(const orc_qpws* w, const double* v, double* out)   /* out[mt] */
{
    const int n = w->n, m = w->m;
    memset(out, 0, sizeof(double) * (size_t)m);
    for (int j = 0; j < n; ++j) {
        double vj = v[j];
        if (vj == 0.0) continue;
        const double* ac = w->A + (size_t)j * m;
        for (int i = 0; i < m; ++i) out[i] += ac[i] * vj;
    }
    for (int j = 0; j < n; ++j) out[m + j] = v[j];
}

/* solve T p = b ; p indexed by absolute Q column (nZ..n-1) */
static void T_solve(const orc_qpws* w, const double* b, double* p)
{
    const int n = w->n, nA = w->nA; const double* Tm = w->T;
    for (int i = 0; i < nA; ++i) {
        int c0 = n - 1 - i;
        double s = b[i];
        for (int c = c0 + 1; c < n; ++c) s -= T_(i, c) * p[c];
        p[c0] = s / T_(i, c0);
    }
}
/* solve T' y = v ; v indexed by absolute column */
static void Tt_solve(const orc_qpws* w, const double* v, double* y)
{
    const int n = w->n, nA = w->nA, nZ = n - nA; const double* Tm = w->T;
    for (int c = nZ; c < n; ++c) {
        int i0 = n - 1 - c;
        double s = v[c];
        for (int i = i0 + 1; i < nA; ++i) s -= T_(i, c) * y[i];
        y[i0] = s / T_(i0, c);
    }
}
/* R' r = v (forward) and R q = r (backward), dimension k */
static void Rt_solve(const orc_qpws* w, int k, const double* v, double* r)
{
    const int n = w->n; const double* Rm = w->R;
    for (int i = 0; i < k; ++i) {
        double s = v[i];
        for (int j = 0; j < i; ++j) s -= R_(j, i) * r[j];
        r[i] = s / R_(i, i);
    }
}
static void R_solve(const orc_qpws* w, int k, const double* r, double* q)
{
    const int n = w->n; const double* Rm = w->R;
    for (int i = k - 1; i >= 0; --i) {
        double s = r[i];
        for (int j = i + 1; j < k; ++j) s -= R_(i, j) * q[j];
        q[i] = s / R_(i, i);
    }
}

/* ------------------------------------------------------------------------------------------
 * Add constraint with row `a` at the end of the working set.  Givens rotations over the nZ
 * null-space columns push a's null-space component into column nZ-1, which becomes the new
 * first range-space column.  Returns 1 (and changes nothing) if a is linearly dependent.
 * ------------------------------------------------------------------------------------------ */
static int add_constraint(orc_qpws* w, int c, int side, const double* a, int updateR)
{
    const int n = w->n, nA = w->nA, nZ = n - nA;
    double* Qm = w->Q; double* Rm = w->R; double* Tm = w->T;
    double* wv = w->tmp4;                    /* Q' a */
    double an = 0, zn = 0;
    for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int i = 0; i < n; ++i) s += Q_(i, j) * a[i];
        wv[j] = s;
    }
    for (int i = 0; i < n; ++i) an += a[i] * a[i];
    for (int j = 0; j < nZ; ++j) zn += wv[j] * wv[j];
    if (nZ == 0 || sqrt(zn) <= EPS_LI * (sqrt(an) > 1.0 ? sqrt(an) : 1.0)) return 1;
    for (int k = 0; k < nZ - 1; ++k) {
        double p = wv[k], q = wv[k + 1];
        double r = hypot(p, q);
        if (r == 0.0) continue;
        double cs = q / r, sn = p / r;
        for (int i = 0; i < n; ++i) {
            double u = Q_(i, k), v = Q_(i, k + 1);
            Q_(i, k) = cs * u - sn * v;
            Q_(i, k + 1) = sn * u + cs * v;
        }
        wv[k] = 0.0; wv[k + 1] = r;
        if (updateR) {
            for (int i = 0; i <= k + 1; ++i) {
                double u = R_(i, k), v = R_(i, k + 1);
                R_(i, k) = cs * u - sn * v;
                R_(i, k + 1) = sn * u + cs * v;
            }
            /* restore triangularity: zero R(k+1,k) with a row rotation */
            double pa = R_(k, k), pb = R_(k + 1, k);
            double rr = hypot(pa, pb);
            if (rr != 0.0) {
                double c2 = pa / rr, s2 = pb / rr;
                for (int j = k; j < nZ; ++j) {
                    double u = R_(k, j), v = R_(k + 1, j);
                    R_(k, j) = c2 * u + s2 * v;
                    R_(k + 1, j) = -s2 * u + c2 * v;
                }
                R_(k + 1, k) = 0.0;
            }
        }
    }
    for (int cc = 0; cc < n; ++cc) T_(nA, cc) = (cc >= nZ - 1) ? wv[cc] : 0.0;
    w->widx[nA] = c;
    w->st[c] = side;
    w->nA = nA + 1;
    return 0;
}

/* Remove working-set position p.  Afterwards Q column (n - nA_old) is the new null-space
 * direction; R is NOT touched (caller appends or refactorises). */
static void remove_constraint(orc_qpws* w, int p)
{
    const int n = w->n, nA = w->nA;
    double* Qm = w->Q; double* Tm = w->T;
    w->st[w->widx[p]] = 0;
    for (int i = p; i < nA - 1; ++i) {
        memcpy(&T_(i, 0), &T_(i + 1, 0), sizeof(double) * (size_t)n);
        w->widx[i] = w->widx[i + 1];
    }
    for (int i = p; i < nA - 1; ++i) {
        int cL = n - 2 - i, cR = n - 1 - i;
        double a = T_(i, cL), b = T_(i, cR);
        double r = hypot(a, b);
        if (r == 0.0) continue;
        double cs = b / r, sn = a / r;
        for (int k = i; k < nA - 1; ++k) {
            double u = T_(k, cL), v = T_(k, cR);
            T_(k, cL) = cs * u - sn * v;
            T_(k, cR) = sn * u + cs * v;
        }
        T_(i, cL) = 0.0;
        for (int k = 0; k < n; ++k) {
            double u = Q_(k, cL), v = Q_(k, cR);
            Q_(k, cL) = cs * u - sn * v;
            Q_(k, cR) = sn * u + cs * v;
        }
    }
    w->nA = nA - 1;
}

/* R = chol(Z'HZ) from scratch; returns 1 if not positive definite */
static int refactor_R(orc_qpws* w)
{
    const int n = w->n, nZ = n - w->nA;
    double* Qm = w->Q; double* Rm = w->R;
    double* hz = w->tmp1;
    w->n_refac++;
    for (int j = 0; j < nZ; ++j) {
        Hmul(w, &Q_(0, j), hz);
        for (int i = 0; i <= j; ++i) {
            double s = 0;
            for (int k = 0; k < n; ++k) s += Q_(k, i) * hz[k];
            R_(i, j) = s;
        }
    }
    for (int j = 0; j < nZ; ++j) {           /* upper Cholesky, column by column */
        for (int i = 0; i < j; ++i) {
            double s = R_(i, j);
            for (int k = 0; k < i; ++k) s -= R_(k, i) * R_(k, j);
            R_(i, j) = s / R_(i, i);
        }
        double d = R_(j, j);
        for (int k = 0; k < j; ++k) d -= R_(k, j) * R_(k, j);
        if (!(d > EPS_PD * w->hscale)) return 1;
        R_(j, j) = sqrt(d);
    }
    return 0;
}

/* KKT step: H dx + dg = A_W' dy,  A_W dx = db   (null-space method) */
static void kkt_step(orc_qpws* w, const double* dg, const double* db, double* dx, double* dyW)
{
    const int n = w->n, nA = w->nA, nZ = n - nA;
    const double* Qm = w->Q;
    double* p = w->tmp1;      /* [n] absolute-column coefficients */
    double* u = w->tmp2;
    double* t = w->tmp3;
    memset(dx, 0, sizeof(double) * (size_t)n);
    if (nA > 0) {
        T_solve(w, db, p);
        for (int c = nZ; c < n; ++c) {
            double pc = p[c];
            if (pc == 0.0) continue;
            for (int i = 0; i < n; ++i) dx[i] += Q_(i, c) * pc;
        }
    }
    Hmul(w, dx, u);
    for (int i = 0; i < n; ++i) u[i] += dg[i];
    if (nZ > 0) {
        for (int j = 0; j < nZ; ++j) {
            double s = 0;
            for (int i = 0; i < n; ++i) s += Q_(i, j) * u[i];
            t[j] = -s;
        }
        Rt_solve(w, nZ, t, p);
        R_solve(w, nZ, p, t);
        double* zq = p;
        memset(zq, 0, sizeof(double) * (size_t)n);
        for (int j = 0; j < nZ; ++j)
            for (int i = 0; i < n; ++i) zq[i] += Q_(i, j) * t[j];
        for (int i = 0; i < n; ++i) dx[i] += zq[i];
        Hmul(w, zq, t);
        for (int i = 0; i < n; ++i) u[i] += t[i];
    }
    if (nA > 0) {
        for (int c = nZ; c < n; ++c) {
            double s = 0;
            for (int i = 0; i < n; ++i) s += Q_(i, c) * u[i];
            t[c] = s;
        }
        Tt_solve(w, t, dyW);
    }
}

/* duals of the working set from scratch, sign-clamped, then re-anchor g_cur = A_W'y - Hx */
static void reanchor(orc_qpws* w, int recompute_y)
{
    const int n = w->n, nA = w->nA, nZ = n - nA;
    const double* Qm = w->Q;
    double* hx = w->tmp1; double* t = w->tmp2; double* yw = w->tmp3; double* a = w->tmp4;
    Hmul(w, w->x, hx);
    if (recompute_y && nA > 0) {
        for (int c = nZ; c < n; ++c) {
            double s = 0;
            for (int i = 0; i < n; ++i) s += Q_(i, c) * (hx[i] + w->g_cur[i]);
            t[c] = s;
        }
        Tt_solve(w, t, yw);
        for (int k = 0; k < nA; ++k) { w->y[w->widx[k]] = yw[k]; if (fabs(yw[k]) > w->ynoise) w->ynoise = fabs(yw[k]); }
    }
    for (int k = 0; k < nA; ++k) {
        int c = w->widx[k];
        int eq = (w->lb_new[c] == w->ub_new[c]);
        if (!eq) {
            if (w->st[c] < 0 && w->y[c] < 0) w->y[c] = 0;
            if (w->st[c] > 0 && w->y[c] > 0) w->y[c] = 0;
        }
    }
    for (int i = 0; i < n; ++i) w->g_cur[i] = -hx[i];
    for (int k = 0; k < nA; ++k) {
        int c = w->widx[k];
        double yc = w->y[c];
        if (yc == 0.0) continue;
        get_row(w, c, a);
        for (int i = 0; i < n; ++i) w->g_cur[i] += a[i] * yc;
    }
}

/* Add constraint `a` when the null space is [PD block of kpd columns (R valid) | one pending
 * column at index kpd].  a must have a non-zero null-space component.  Afterwards the null space
 * is [PD block of kpd-1 columns | pending column at index kpd-1] (or empty when kpd == 0). */
static void add_constraint_pending(orc_qpws* w, int c, int side, const double* a, int kpd)
{
    const int n = w->n, nA = w->nA;
    double* Qm = w->Q; double* Rm = w->R; double* Tm = w->T;
    double* wv = w->tmp4;
    for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int i = 0; i < n; ++i) s += Q_(i, j) * a[i];
        wv[j] = s;
    }
    for (int k = 0; k < kpd; ++k) {
        double p = wv[k], q = wv[k + 1];
        double r = hypot(p, q);
        if (r == 0.0) continue;
        double cs = q / r, sn = p / r;
        for (int i = 0; i < n; ++i) {
            double u = Q_(i, k), v = Q_(i, k + 1);
            Q_(i, k) = cs * u - sn * v;
            Q_(i, k + 1) = sn * u + cs * v;
        }
        wv[k] = 0.0; wv[k + 1] = r;
        if (k < kpd - 1) {                       /* both columns inside the PD block: update R */
            for (int i = 0; i <= k + 1; ++i) {
                double u = R_(i, k), v = R_(i, k + 1);
                R_(i, k) = cs * u - sn * v;
                R_(i, k + 1) = sn * u + cs * v;
            }
            double pa = R_(k, k), pb = R_(k + 1, k);
            double rr = hypot(pa, pb);
            if (rr != 0.0) {
                double c2 = pa / rr, s2 = pb / rr;
                for (int j = k; j < kpd; ++j) {
                    double u = R_(k, j), v = R_(k + 1, j);
                    R_(k, j) = c2 * u + s2 * v;
                    R_(k + 1, j) = -s2 * u + c2 * v;
                }
                R_(k + 1, k) = 0.0;
            }
        }
    }
    for (int cc = 0; cc < n; ++cc) T_(nA, cc) = (cc >= kpd) ? wv[cc] : 0.0;
    w->widx[nA] = c;
    w->st[c] = side;
    w->nA = nA + 1;
}

static void reanchor(orc_qpws* w, int recompute_y);

/* The null space is [PD block (kpd = n-nA-1 columns, R valid) | pending column].  Either append
 * the pending column to R (projected Hessian stays positive definite) or move along the
 * direction of non-positive curvature to the next blocking constraint, add it and repeat.
 * (c_out, side_out) = the constraint whose removal created the pending column. */
static int resolve_pending(orc_qpws* w, int c_out, int side_out, int dbg, int zero_step)
{
    const int n = w->n, mt = w->mt;
    double* Qm = w->Q; double* Rm = w->R;
    int first = 1, moved = 0;
    for (;;) {
        const int kpd = n - w->nA - 1;
        if (kpd < 0) break;                              /* null space empty */
        double* z = &Q_(0, kpd);
        double* hz = w->tmp1; double* v = w->tmp2; double* r = w->tmp3;
        Hmul(w, z, hz);
        double zhz = 0;
        for (int i = 0; i < n; ++i) zhz += z[i] * hz[i];
        for (int j = 0; j < kpd; ++j) {
            double s = 0;
            for (int i = 0; i < n; ++i) s += Q_(i, j) * hz[i];
            v[j] = s;
        }
        Rt_solve(w, kpd, v, r);
        double rho2 = zhz;
        for (int j = 0; j < kpd; ++j) rho2 -= r[j] * r[j];
        if (rho2 > EPS_PD * w->hscale) {
            for (int j = 0; j < kpd; ++j) R_(j, kpd) = r[j];
            R_(kpd, kpd) = sqrt(rho2);
            break;
        }
        /* direction of non-positive curvature, H-conjugate to the PD block */
        double* q = w->tmp4; double* d = w->dx;
        R_solve(w, kpd, r, q);
        for (int i = 0; i < n; ++i) d[i] = z[i];
        for (int j = 0; j < kpd; ++j)
            for (int i = 0; i < n; ++i) d[i] -= Q_(i, j) * q[j];
        double sgn = 0;
        if (first) {                                     /* move off the bound that was left */
            double a_c[ORC_NMAX];
            get_row(w, c_out, a_c);
            for (int i = 0; i < n; ++i) sgn += a_c[i] * d[i];
            if (side_out > 0) sgn = -sgn;
        } else {                                         /* descent direction of the aux QP */
            double gn = 0, dn2 = 0;
            Hmul(w, w->x, hz);
            for (int i = 0; i < n; ++i) { double gi = hz[i] + w->g_cur[i]; sgn -= gi * d[i]; gn += gi * gi; dn2 += d[i] * d[i]; }
            if (fabs(sgn) <= FLAT_TOL * sqrt(gn) * sqrt(dn2)) {
                /* the gradient is orthogonal to d within round-off (flat direction): both senses are
                 * descent directions of equal merit; take a canonical one so that every arithmetic
                 * makes the same choice: away from the bound that was left, else first significant
                 * component of d positive */
                double a_c[ORC_NMAX], an = 0;
                get_row(w, c_out, a_c);
                sgn = 0;
                for (int i = 0; i < n; ++i) { sgn += a_c[i] * d[i]; an += a_c[i] * a_c[i]; }
                if (side_out > 0) sgn = -sgn;
                if (fabs(sgn) <= FLAT_TOL * sqrt(an) * sqrt(dn2)) {
                    double dmx = 0;
                    for (int i = 0; i < n; ++i) if (fabs(d[i]) > dmx) dmx = fabs(d[i]);
                    sgn = 1.0;
                    for (int i = 0; i < n; ++i) if (fabs(d[i]) > 0.25 * dmx) { sgn = d[i]; break; }
                }
            }
        }
        if (dbg) fprintf(stderr, "   exch-dir: first %d sgn %.3e\n", first, sgn);
        if (sgn < 0) for (int i = 0; i < n; ++i) d[i] = -d[i];
        Amul(w, d, w->Adx);
        double tmin = INFINITY; int j = -1, jside = 0;
        double dn = 0; for (int i = 0; i < n; ++i) dn += d[i] * d[i]; dn = sqrt(dn);
        double xmax = 0; for (int i = 0; i < n; ++i) if (fabs(w->x[i]) > xmax) xmax = fabs(w->x[i]);
        rt_cand* cd = w->cand; int nc = 0;
        for (int cc = 0; cc < mt; ++cc) {
            if (w->st[cc] != 0 || w->anorm[cc] == 0.0) continue;
            double adx = w->Adx[cc];
            const double rz = EPS_ROW * w->anorm[cc] * dn;
            const double del = w->tie * w->anorm[cc] * (1.0 + xmax);
            if (adx < -rz && !isinf(w->lb_cur[cc])) {
                double sl = w->Ax[cc] - w->lb_cur[cc]; if (sl < 0) sl = 0;
                cd[nc].num = sl; cd[nc].den = -adx; cd[nc].del = del; cd[nc].key = 2 * cc; cd[nc].aux = cc; ++nc;
            } else if (adx > rz && !isinf(w->ub_cur[cc])) {
                double sl = w->ub_cur[cc] - w->Ax[cc]; if (sl < 0) sl = 0;
                cd[nc].num = sl; cd[nc].den = adx; cd[nc].del = del; cd[nc].key = 2 * cc + 1; cd[nc].aux = cc; ++nc;
            }
        }
        {
            const int sel = rt_select(cd, nc);
            if (sel >= 0) { tmin = (cd[sel].num <= cd[sel].del) ? 0.0 : cd[sel].num / cd[sel].den; j = cd[sel].aux; jside = (cd[sel].key & 1) ? 1 : -1; }
        }
        if (dbg) fprintf(stderr, "   exch: rho2 %.3e zhz %.3e out %d -> j %d jside %d t %.3e kpd %d\n", rho2, zhz, c_out, j, jside, tmin, kpd);
        if (j < 0) return EEPACC_QP_UNBOUNDED;
        for (int i = 0; i < n; ++i) w->x[i] += tmin * d[i];
        for (int cc = 0; cc < mt; ++cc) w->Ax[cc] += tmin * w->Adx[cc];
        if (jside < 0) w->lb_cur[j] = w->Ax[j]; else w->ub_cur[j] = w->Ax[j];
        double a_j[ORC_NMAX];
        get_row(w, j, a_j);
        add_constraint_pending(w, j, jside, a_j, kpd);
        w->y[j] = 0.0;
        if (zero_step) w->tabu[j] = 1;
        w->n_exch++;
        first = 0; moved = 1;
    }
    if (moved) reanchor(w, 1);
    return 0;
}

static void cold_ws(orc_qpws* w)
{
    const int n = w->n, m = w->m;
    double* Qm = w->Q; double* Tm = w->T;
    memset(w->st, 0, sizeof(int) * (size_t)w->mt);
    memset(w->y, 0, sizeof(double) * (size_t)w->mt);
    memset(Qm, 0, sizeof(double) * (size_t)n * n);
    memset(Tm, 0, sizeof(double) * (size_t)n * n);
    for (int i = 0; i < n; ++i) Q_(i, i) = 1.0;
    for (int i = 0; i < n; ++i) {       /* position i holds variable n-1-i */
        int v = n - 1 - i;
        w->widx[i] = m + v;
        w->st[m + v] = -1;
        /* distinct positive initial duals break the all-zero tie (cf. qpOASES "ramping") */
        w->y[m + v] = 0.5 + 0.5 * (double)(m + v) / (double)w->mt;
        T_(i, v) = 1.0;
    }
    w->nA = n;
}

/* auxiliary QP for which the current (x, y, working set) is optimal: relaxed bounds + anchored gradient */
static void setup_aux(orc_qpws* w)
{
    const int mt = w->mt;
    Amul(w, w->x, w->Ax);
    for (int c = 0; c < mt; ++c) {
        int s = w->st[c];
        double ax = w->Ax[c];
        int eq = (w->lb_new[c] == w->ub_new[c]);
        if (s < 0) {
            w->lb_cur[c] = ax;
            w->ub_cur[c] = eq ? ax : (isinf(w->ub_new[c]) ? INFINITY : ax + RELAX);
        } else if (s > 0) {
            w->ub_cur[c] = ax;
            w->lb_cur[c] = eq ? ax : (isinf(w->lb_new[c]) ? -INFINITY : ax - RELAX);
        } else {
            w->lb_cur[c] = isinf(w->lb_new[c]) ? -INFINITY : ax - RELAX;
            w->ub_cur[c] = isinf(w->ub_new[c]) ? INFINITY : ax + RELAX;
        }
    }
    reanchor(w, 0);
}

int orc_qpsolve(orc_qpws* w, const double* H, const double* g, const double* A,
                const double* lba, const double* uba, const double* lbx, const double* ubx,
                double* xout, double* cost, double* lam_a, double* lam_x, int* iter_out)
{
    const int n = w->n, m = w->m, mt = w->mt;
    int status = EEPACC_QP_OK, iter = 0;
    const int max_iter = 5 * (n + m);
    const int restart_after = (4 * n > 400 ? 4 * n : 400);
    w->H = H; w->A = A; w->g_new = g;
    w->n_exch = 0; w->n_refac = 0;
    memset(w->tabu, 0, (size_t)w->mt);
    const int dbg = getenv("ORC_QP_DEBUG") != 0;
    const int reanchor_left = getenv("ORC_QP_REANCHOR_LEFT") != 0;
    for (int i = 0; i < m; ++i) {
        w->lb_new[i] = (lba[i] <= -BIG) ? -INFINITY : lba[i];
        w->ub_new[i] = (uba[i] >= BIG) ? INFINITY : uba[i];
    }
    for (int j = 0; j < n; ++j) {
        double l = lbx ? lbx[j] : -INFINITY, u = ubx ? ubx[j] : INFINITY;
        w->lb_new[m + j] = (l < -FARB) ? -FARB : l;
        w->ub_new[m + j] = (u > FARB) ? FARB : u;
    }
    w->hscale = 1e-300;
    for (int j = 0; j < n; ++j) { double d = fabs(H[(size_t)j * n + j]); if (d > w->hscale) w->hscale = d; }
    /* H == 0 (the BL/TV LPs, Settings.m:65-76): qpOASES treats a zero Hessian as eps*I
     * (HST_ZERO + regularisation), which selects the least-norm LP optimum; for eps below a
     * problem-dependent threshold that point does not depend on eps (Mangasarian & Meyer 1979). */
    w->reg = 0.0;
    {
        double hmax = 0;
        for (size_t k = 0; k < (size_t)n * n; ++k) if (fabs(H[k]) > hmax) hmax = fabs(H[k]);
        if (hmax == 0.0) w->reg = LP_REG;
    }
    if (w->hscale < LP_REG) w->hscale = LP_REG;
    /* tie tolerance of the ratio tests.  The eps-regularised LPs keep exact comparisons: with a curvature of
     * 1e-6 against multipliers of 1e7 a tie tolerance on the duals would have to be below round-off to leave
     * the least-norm optimum alone, and their optimum is unique anyway (no flat directions). */
    w->tie = (w->reg != 0.0) ? 0.0 : TIE_TOL;

    for (int c = 0; c < m; ++c) w->anorm[c] = 0.0;
    for (int j = 0; j < n; ++j)
        for (int c = 0; c < m; ++c) w->anorm[c] += fabs(A[(size_t)j * m + c]);
    for (int j = 0; j < n; ++j) w->anorm[m + j] = 1.0;

    /* ---------------- set-up: working set, factorisation, auxiliary QP ---------------- */
    int need_cold = !w->warm;
    if (!need_cold) {
        /* rebuild TQ for the kept working set under the new A; drop dependent / unbounded rows */
        int old_nA = w->nA;
        int* oldidx = (int*)malloc(sizeof(int) * (size_t)(old_nA ? old_nA : 1));
        int* oldst = (int*)malloc(sizeof(int) * (size_t)(old_nA ? old_nA : 1));
        for (int k = 0; k < old_nA; ++k) { oldidx[k] = w->widx[k]; oldst[k] = w->st[w->widx[k]]; }
        memset(w->st, 0, sizeof(int) * (size_t)mt);
        double* Qm = w->Q;
        memset(Qm, 0, sizeof(double) * (size_t)n * n);
        for (int i = 0; i < n; ++i) Q_(i, i) = 1.0;
        w->nA = 0;
        double* a = w->tmp3;
        for (int k = 0; k < old_nA; ++k) {
            int c = oldidx[k], s = oldst[k];
            double bnd = s < 0 ? w->lb_new[c] : w->ub_new[c];
            if (isinf(bnd)) { w->y[c] = 0; continue; }
            get_row(w, c, a);
            if (add_constraint(w, c, s, a, 0)) w->y[c] = 0;
        }
        free(oldidx); free(oldst);
        if (refactor_R(w)) need_cold = 1;
    }
    if (need_cold) {
        if (!w->warm) memset(w->x, 0, sizeof(double) * (size_t)n);
        cold_ws(w);
    }
    setup_aux(w);
    w->warm = 1;
    w->ynoise = 1.0;
    for (int k = 0; k < w->nA; ++k) if (fabs(w->y[w->widx[k]]) > w->ynoise) w->ynoise = fabs(w->y[w->widx[k]]);

    /* ---------------- homotopy ---------------- */
    double* db = w->tmp4;
    for (;;) {
        const int nA = w->nA;
        for (int i = 0; i < n; ++i) w->dg[i] = g[i] - w->g_cur[i];
        for (int c = 0; c < mt; ++c) {
            w->dlb[c] = isinf(w->lb_new[c]) ? 0.0 : w->lb_new[c] - w->lb_cur[c];
            w->dub[c] = isinf(w->ub_new[c]) ? 0.0 : w->ub_new[c] - w->ub_cur[c];
        }
        double dbw[ORC_NMAX];
        for (int k = 0; k < nA; ++k) {
            int c = w->widx[k];
            dbw[k] = w->st[c] < 0 ? w->dlb[c] : w->dub[c];
        }
        (void)db;
        kkt_step(w, w->dg, dbw, w->dx, w->dyW);
        Amul(w, w->dx, w->Adx);

        /* ratio tests (two-pass, see rt_select): primal candidates keyed 2c+side, dual candidates 2mt+c */
        double tau = 1.0; int blk = -1, blk_side = 0, blk_dual = -1;
        double dxmax = 0, xmax = 0, ymax = 0;
        for (int i = 0; i < n; ++i) { if (fabs(w->dx[i]) > dxmax) dxmax = fabs(w->dx[i]); if (fabs(w->x[i]) > xmax) xmax = fabs(w->x[i]); }
        for (int k = 0; k < nA; ++k) if (fabs(w->y[w->widx[k]]) > ymax) ymax = fabs(w->y[w->widx[k]]);
        rt_cand* cd = w->cand; int nc = 0;
        for (int c = 0; c < mt; ++c) {
            if (w->st[c] != 0 || w->anorm[c] == 0.0) continue;     /* all-zero rows can never be added */
            if (w->tabu[c] == 2) continue;                         /* left during this run of zero steps: no re-entry */
            double adx = w->Adx[c];
            const double rz = EPS_ROW * w->anorm[c] * dxmax;
            const double del = w->tie * w->anorm[c] * (1.0 + xmax);
            if (!isinf(w->lb_cur[c])) {
                double rate = adx - w->dlb[c];
                double sl = w->Ax[c] - w->lb_cur[c]; if (sl < 0) sl = 0;
                if (rate < -(EPS_RATE * (1.0 + fabs(adx) + fabs(w->dlb[c])) + rz) &&
                    -rate - sl > FEAS_TOL * (1.0 + fabs(w->lb_new[c]))) {
                    if (sl < -rate) { cd[nc].num = sl; cd[nc].den = -rate; cd[nc].del = del; cd[nc].key = 2 * c; cd[nc].aux = c; ++nc; }
                }
            }
            if (!isinf(w->ub_cur[c])) {
                double rate = w->dub[c] - adx;
                double sl = w->ub_cur[c] - w->Ax[c]; if (sl < 0) sl = 0;
                if (rate < -(EPS_RATE * (1.0 + fabs(adx) + fabs(w->dub[c])) + rz) &&
                    -rate - sl > FEAS_TOL * (1.0 + fabs(w->ub_new[c]))) {
                    if (sl < -rate) { cd[nc].num = sl; cd[nc].den = -rate; cd[nc].del = del; cd[nc].key = 2 * c + 1; cd[nc].aux = c; ++nc; }
                }
            }
        }
        double dys = 0;
        for (int k = 0; k < nA; ++k) if (fabs(w->dyW[k]) > dys) dys = fabs(w->dyW[k]);
        for (int k = 0; k < nA; ++k) {
            int c = w->widx[k];
            if (w->lb_new[c] == w->ub_new[c]) continue;          /* equality: never leaves */
            if (w->tabu[c] == 1) continue;                       /* entered during this run of zero steps: stays */
            double yc = w->y[c], dy = w->dyW[k];
            double thr = EPS_RATE * (1.0 + dys);
            const double ysc = 1.0 + fabs(yc) + fabs(dy);         /* scale of this multiplier along the step */
            const double del = w->tie * w->ynoise;
            /* a dual that would end within dtol of zero at the full step is not blocking (dual twin of FEAS_TOL) */
            const double dtol = fmax(FEAS_TOL * ysc, del);
            if (w->st[c] < 0 && dy < -thr) {
                double num = (yc > 0 ? yc : 0);
                if (num < -dy && -dy - num > dtol) { cd[nc].num = num; cd[nc].den = -dy; cd[nc].del = del; cd[nc].key = 2 * mt + c; cd[nc].aux = k; ++nc; }
            } else if (w->st[c] > 0 && dy > thr) {
                double num = (yc < 0 ? -yc : 0);
                if (num < dy && dy - num > dtol) { cd[nc].num = num; cd[nc].den = dy; cd[nc].del = del; cd[nc].key = 2 * mt + c; cd[nc].aux = k; ++nc; }
            }
        }
        {
            const int sel = rt_select(cd, nc);
            if (sel >= 0) {
                tau = (cd[sel].num <= cd[sel].del) ? 0.0 : cd[sel].num / cd[sel].den;   /* within round-off of zero: a zero step */
                if (cd[sel].key < 2 * mt) { blk = cd[sel].aux; blk_side = (cd[sel].key & 1) ? 1 : -1; }
                else blk_dual = cd[sel].aux;
            }
        }
        /* anti-cycling: constraints that enter while the homotopy makes no progress may not leave
         * again through the dual test until it does (flat edges at degenerate vertices otherwise
         * ping-pong between two equally optimal vertices) */
        const int zero_step = !(tau > 1e-14);
        if (!zero_step) memset(w->tabu, 0, (size_t)mt);

        /* take the step */
        w->ynoise += tau * dys;
        for (int i = 0; i < n; ++i) { w->x[i] += tau * w->dx[i]; }
        for (int k = 0; k < nA; ++k) w->y[w->widx[k]] += tau * w->dyW[k];
        for (int c = 0; c < mt; ++c) w->Ax[c] += tau * w->Adx[c];
        if (tau >= 1.0) {
            for (int i = 0; i < n; ++i) w->g_cur[i] = g[i];
            for (int c = 0; c < mt; ++c) {
                if (!isinf(w->lb_new[c])) w->lb_cur[c] = w->lb_new[c];
                if (!isinf(w->ub_new[c])) w->ub_cur[c] = w->ub_new[c];
            }
            break;
        }
        for (int i = 0; i < n; ++i) w->g_cur[i] += tau * w->dg[i];
        for (int c = 0; c < mt; ++c) { w->lb_cur[c] += tau * w->dlb[c]; w->ub_cur[c] += tau * w->dub[c]; }
        if (++iter > max_iter) { status = EEPACC_QP_MAXITER; break; }
        if (iter == restart_after) {
            /* the warm-started path wanders (degenerate / non-convex stretches): start over from an
             * all-fixed working set at the current point -- the cold-start path */
            cold_ws(w);
            setup_aux(w);
            memset(w->tabu, 0, (size_t)mt);
            continue;
        }
        if (dbg) {   /* stationarity of the homotopy iterate: H x + g_cur - A_W' y */
            double hx[ORC_NMAX], a_[ORC_NMAX], rs = 0;
            Hmul(w, w->x, hx);
            for (int i = 0; i < n; ++i) hx[i] += w->g_cur[i];
            for (int k = 0; k < w->nA; ++k) { int c = w->widx[k]; get_row(w, c, a_); for (int i = 0; i < n; ++i) hx[i] -= a_[i] * w->y[c]; }
            for (int i = 0; i < n; ++i) if (fabs(hx[i]) > rs) rs = fabs(hx[i]);
            fprintf(stderr, "   stationarity of the iterate %.3e\n", rs);
        }
        if (dbg) fprintf(stderr, "it %d tau %.3e blk %d side %d dual %d (c=%d) nA %d dy %.3e dys %.3e y %.3e dxmax %.3e\n", iter, tau, blk, blk_side, blk_dual, blk_dual >= 0 ? w->widx[blk_dual] : -1, w->nA, blk_dual >= 0 ? w->dyW[blk_dual] : 0.0, dys, blk_dual >= 0 ? w->y[w->widx[blk_dual]] : 0.0, dxmax);

        if (blk >= 0) {
            /* ---------- primal blocking: add constraint blk ---------- */
            double* a = w->tmp3;
            get_row(w, blk, a);
            /* snap the value to the bound it hit */
            if (blk_side < 0) w->lb_cur[blk] = w->Ax[blk]; else w->ub_cur[blk] = w->Ax[blk];
            if (zero_step) w->tabu[blk] = 1;
            if (add_constraint(w, blk, blk_side, a, 1) == 0) {
                w->y[blk] = 0.0;
            } else {
                /* linearly dependent: a = A_W' lam ; exchange by dual ratio test */
                const double* Qm = w->Q;
                const int nZ = n - w->nA;
                double* t = w->tmp1; double* lam = w->tmp2;
                for (int c = nZ; c < n; ++c) {
                    double s = 0;
                    for (int i = 0; i < n; ++i) s += Q_(i, c) * a[i];
                    t[c] = s;
                }
                Tt_solve(w, t, lam);
                double lmax = 0;
                for (int k = 0; k < w->nA; ++k) if (fabs(lam[k]) > lmax) lmax = fabs(lam[k]);
                double theta = INFINITY; int kout = -1;
                double ymx = 0;
                for (int k = 0; k < w->nA; ++k) if (fabs(w->y[w->widx[k]]) > ymx) ymx = fabs(w->y[w->widx[k]]);
                rt_cand* cd2 = w->cand; int nc2 = 0;
                for (int k = 0; k < w->nA; ++k) {
                    int c = w->widx[k];
                    if (w->lb_new[c] == w->ub_new[c]) continue;
                    double l = (blk_side < 0) ? lam[k] : -lam[k];   /* y_k(theta) = y_k - theta*l */
                    if (fabs(l) <= 1e-11 * (1.0 + lmax)) continue;
                    double yc = w->y[c];
                    if (w->st[c] < 0 && l > 0) {
                        cd2[nc2].num = (yc > 0 ? yc : 0); cd2[nc2].den = l;
                    } else if (w->st[c] > 0 && l < 0) {
                        cd2[nc2].num = (yc < 0 ? -yc : 0); cd2[nc2].den = -l;
                    } else continue;
                    cd2[nc2].del = w->tie * w->ynoise; cd2[nc2].key = c; cd2[nc2].aux = k; ++nc2;
                }
                {
                    const int sel = rt_select(cd2, nc2);
                    if (sel >= 0) { theta = (cd2[sel].num <= cd2[sel].del) ? 0.0 : cd2[sel].num / cd2[sel].den; kout = cd2[sel].aux; }
                }
                if (kout < 0) { status = EEPACC_QP_INFEASIBLE; break; }
                w->ynoise += theta * lmax;
                for (int k = 0; k < w->nA; ++k) {
                    double l = (blk_side < 0) ? lam[k] : -lam[k];
                    w->y[w->widx[k]] -= theta * l;
                }
                int cout_ = w->widx[kout];
                w->y[cout_] = 0.0;
                if (zero_step) w->tabu[cout_] = 2;
                remove_constraint(w, kout);
                {   /* new T row for blk over the (unchanged-size) range space; R untouched */
                    double* Tm = w->T;
                    const int nA2 = w->nA, nZ2 = n - nA2 - 1;     /* freed column = n-nA2-1 */
                    for (int cc = 0; cc < n; ++cc) {
                        double s = 0;
                        if (cc >= nZ2) for (int i = 0; i < n; ++i) s += Q_(i, cc) * a[i];
                        T_(nA2, cc) = s;
                    }
                    w->widx[nA2] = blk; w->st[blk] = blk_side; w->nA = nA2 + 1;
                }
                w->y[blk] = (blk_side < 0) ? theta : -theta;
                w->n_exch++;
            }
        } else if (blk_dual >= 0) {
            /* ---------- dual blocking: remove working-set position blk_dual ---------- */
            const int c = w->widx[blk_dual];
            const int side = w->st[c];
            const double y_left = w->y[c];
            w->y[c] = 0.0;
            if (zero_step) w->tabu[c] = 2;
            remove_constraint(w, blk_dual);
            int rc = resolve_pending(w, c, side, dbg, zero_step);
            if (rc) { status = rc; break; }
            /* A multiplier normally leaves at zero (to round-off, or within the tie tolerance of a zero step).  One that was
             * shielded by the anti-cycling flag while the homotopy advanced leaves with a finite, wrong-signed value: dropping
             * it would leave the iterate non-stationary for the current homotopy data for the rest of the solve (found by the
             * interior-point cross-check: 6 of 55 744 ABMPC bench QPs ended on a non-optimal vertex).  Re-anchoring the gradient
             * homotopy at the iterate instead removes all six (scripts/cpu_oracle_kkt_audit.py with ORC_QP_REANCHOR_LEFT=1) and
             * leaves the golden runs as they are.  OFF by default: the rule fires about once per 1000 QPs, so the GPU kernels
             * (k3r_qp.cu, k3_qp.cu) have to carry it too before the default can change -- DESIGN.md section 7. */
            if (reanchor_left && fabs(y_left) > fmax(w->tie, 1e-9) * w->ynoise) reanchor(w, 0);
        }
    }

    /* ---------------- iterative refinement on the final working set ---------------- */
    if (status == EEPACC_QP_OK && !getenv("ORC_NO_REFINE")) {
        for (int pass = 0; pass < 1; ++pass) {
            const int nA = w->nA;
            double* hx = w->tmp1; double* a = w->tmp3;
            double rg[ORC_NMAX], rb[ORC_NMAX];
            Hmul(w, w->x, hx);
            for (int i = 0; i < n; ++i) rg[i] = hx[i] + g[i];
            for (int k = 0; k < nA; ++k) {
                int c = w->widx[k];
                get_row(w, c, a);
                double ax = 0, yc = w->y[c];
                for (int i = 0; i < n; ++i) { ax += a[i] * w->x[i]; rg[i] -= a[i] * yc; }
                rb[k] = (w->st[c] < 0 ? w->lb_cur[c] : w->ub_cur[c]) - ax;
            }
            kkt_step(w, rg, rb, w->dx, w->dyW);
            for (int i = 0; i < n; ++i) w->x[i] += w->dx[i];
            for (int k = 0; k < nA; ++k) w->y[w->widx[k]] += w->dyW[k];
        }
    }

    /* ---------------- outputs ---------------- */
    memcpy(xout, w->x, sizeof(double) * (size_t)n);
    if (cost) {
        double* hx = w->tmp1;
        Hmul(w, w->x, hx);
        double s = 0;
        for (int i = 0; i < n; ++i) s += w->x[i] * (0.5 * (hx[i] - w->reg * w->x[i]) + g[i]);
        *cost = s;
    }
    /* CasADi sign convention (multiplier > 0 at an upper bound) = -qpOASES' */
    if (lam_a) for (int i = 0; i < m; ++i) lam_a[i] = -w->y[i];
    if (lam_x) for (int j = 0; j < n; ++j) lam_x[j] = -w->y[m + j];
    if (iter_out) *iter_out = iter;
    if (status == EEPACC_QP_OK) {
        for (int j = 0; j < n; ++j)
            if (w->st[m + j] != 0 && fabs(w->x[j]) >= FARB) status = EEPACC_QP_UNBOUNDED;
        for (int c = 0; c < m; ++c)       /* an all-zero row with inconsistent bounds: infeasible QP */
            if (w->anorm[c] == 0.0 && (w->lb_new[c] > 1e-8 || w->ub_new[c] < -1e-8)) status = EEPACC_QP_INFEASIBLE;
    }
    if (status != EEPACC_QP_OK) w->warm = 0;
    return status;
}
