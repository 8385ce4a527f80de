// k7_nlp.cuh -- K7: the full-route NLP of MATLAB_CasADi/RunOpt_NLP.m:204-510, one scenario per thread.
//
// The reference hands the problem (8 700 variables for the 870-interval golden cycle) to IPOPT through CasADi.
// IPOPT's source is not in the tree (binary-only CasADi plugin), so what is kept is the PROBLEM, statement by
// statement (stage function k7_stage below, every row citing RunOpt_NLP.m), and the published method class:
// a primal-dual interior-point method with a monotone barrier schedule, fraction-to-the-boundary steps and a
// backtracking line search (Waechter & Biegler 2006, sections 2-3), restated for the optimal-control structure
// the problem has:
//   * the Newton system is solved by a Riccati sweep over the stages (state 4, control 6, <= 36 inequality rows
//     per stage) instead of a sparse LDL' of the whole KKT matrix -- O(N) work and O(N) memory per scenario,
//     which is what lets one GPU thread own one scenario;
//   * the step is applied as a closed-loop nonlinear roll-out (u + alpha k + K dx), so the shooting gaps of
//     RunOpt_NLP.m:372-375 are zero at every iterate and the slack update t+ = (1-alpha)(d + t) - d(z+) makes the
//     inequality residual shrink by exactly (1 - alpha);
//   * derivatives come from forward-mode dual numbers with 10 directions (x, u) -- the same device as the K1
//     evaluator's F_jac -- and the Lagrangian Hessian is the stage-cost Hessian plus the multiplier-weighted
//     curvature of the drag, motor-power and headway-policy terms; it is convexified by a diagonal shift whenever a
//     stage's reduced Hessian fails its Cholesky factorisation.
// The state carried across stages is (s, v, c, j) with c = Fm + Fb - zeta v^2 - m g (c_r cos + sin)(theta) of the
// previous interval, which turns the jerk definition RunOpt_NLP.m:383-386 into a state equation; theta is a function of
// s (:377-382) and is substituted.  The optimiser of this problem is the optimiser of the reference's.
//
// Compiles for the device (k7_nlp.cu) and, unchanged, for the host (tests/host_emu/k7_emu.cpp: TEST INFRASTRUCTURE).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "../../include/eepacc.h"
#include "../../include/eepacc_nlp.h"

#define K7_HD __host__ __device__ __forceinline__
#define K7_NX 4
#define K7_NU 6
#define K7_NZ 10
#define K7_ROWS_BASE 28
#define K7_MAXR (K7_ROWS_BASE + 2 * EEPACC_MAX_TL)
#define K7_PI 3.14159265358979323846
#ifndef K7_KAPPA_MU
#define K7_KAPPA_MU 0.2      // barrier schedule mu+ = max(tol / 10, min(kappa mu, mu^theta))
#define K7_THETA_MU 1.5
#endif

// ---- per-stage workspace fields (doubles), workspace element (stage k, field f) = ws[(k * K7_STAGE + f) * stride]
#define K7_F_X    0
#define K7_F_U    (K7_F_X + K7_NX)
#define K7_F_T    (K7_F_U + K7_NU)
#define K7_F_LAM  (K7_F_T + K7_MAXR)
#define K7_F_RP   (K7_F_LAM + K7_MAXR)
#define K7_F_KFF  (K7_F_RP + K7_MAXR)
#define K7_F_KFB  (K7_F_KFF + K7_NU)
#define K7_F_XT   (K7_F_KFB + K7_NU * K7_NX)
#define K7_F_UT   (K7_F_XT + K7_NX)
#define K7_F_TT   (K7_F_UT + K7_NU)
#define K7_F_GL   (K7_F_TT + K7_MAXR)
#define K7_F_L0   (K7_F_GL + K7_NZ)
#define K7_STAGE  (K7_F_L0 + 1)

// ------------------------------------------------------------------------------------------------------------
// forward-mode dual numbers, 10 directions
// ------------------------------------------------------------------------------------------------------------
template <int ND> struct K7DualT { double v; double d[ND]; };
typedef K7DualT<K7_NZ> K7Dual;       // all 10 directions (x, u): Jacobians for the backward sweep

#define K7_T template <int ND> K7_HD
K7_T K7DualT<ND> operator+(const K7DualT<ND>& a, const K7DualT<ND>& b) { K7DualT<ND> r; r.v = a.v + b.v; for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
K7_T K7DualT<ND> operator-(const K7DualT<ND>& a, const K7DualT<ND>& b) { K7DualT<ND> r; r.v = a.v - b.v; for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
K7_T K7DualT<ND> operator*(const K7DualT<ND>& a, const K7DualT<ND>& b) { K7DualT<ND> r; r.v = a.v * b.v; for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
K7_T K7DualT<ND> operator+(const K7DualT<ND>& a, double b) { K7DualT<ND> r = a; r.v += b; return r; }
K7_T K7DualT<ND> operator-(const K7DualT<ND>& a, double b) { K7DualT<ND> r = a; r.v -= b; return r; }
K7_T K7DualT<ND> operator*(const K7DualT<ND>& a, double b) { K7DualT<ND> r; r.v = a.v * b; for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] * b; return r; }
K7_T K7DualT<ND> operator*(double b, const K7DualT<ND>& a) { return a * b; }
K7_T K7DualT<ND> operator+(double b, const K7DualT<ND>& a) { return a + b; }
K7_T K7DualT<ND> operator-(double b, const K7DualT<ND>& a) { K7DualT<ND> r; r.v = b - a.v; for (int i = 0; i < ND; ++i) r.d[i] = -a.d[i]; return r; }
K7_T K7DualT<ND> operator-(const K7DualT<ND>& a) { K7DualT<ND> r; r.v = -a.v; for (int i = 0; i < ND; ++i) r.d[i] = -a.d[i]; return r; }

K7_HD double k7_val(double a) { return a; }
K7_T double k7_val(const K7DualT<ND>& a) { return a.v; }
// g(a) with derivative dg: chain rule
K7_HD double k7_chain(double, double g, double) { return g; }
K7_T K7DualT<ND> k7_chain(const K7DualT<ND>& a, double g, double dg) { K7DualT<ND> r; r.v = g; for (int i = 0; i < ND; ++i) r.d[i] = dg * a.d[i]; return r; }
template <class T> K7_HD T k7_cos(const T& a) { const double x = k7_val(a); return k7_chain(a, cos(x), -sin(x)); }
template <class T> K7_HD T k7_sin(const T& a) { const double x = k7_val(a); return k7_chain(a, sin(x), cos(x)); }
K7_HD void k7_const(double& a, double c) { a = c; }
K7_T void k7_const(K7DualT<ND>& a, double c) { a.v = c; for (int i = 0; i < ND; ++i) a.d[i] = 0.0; }

// casadi.interpolant('LUT','linear',{xs},ys): piece-wise linear, linear extrapolation from the end cells
// (RunOpt_NLP.m:66-112,155,184).  A grid of one point is a constant.
template <class T> K7_HD T k7_lut(const T& s, const double* xs, const double* ys, int n)
{
    const double x = k7_val(s);
    if (n <= 1) return k7_chain(s, n == 1 ? ys[0] : 0.0, 0.0);
    int i = 0;
    while (i < n - 2 && x >= xs[i + 1]) ++i;
    const double dx = xs[i + 1] - xs[i];
    const double sl = dx > 0.0 ? (ys[i + 1] - ys[i]) / dx : 0.0;
    return k7_chain(s, ys[i] + sl * (x - xs[i]), sl);
}

// alpha_TTL * |kappa|^(-1/3) (RunOpt_NLP.m:455); kappa = 0 would be an infinite speed: capped at 1e5 like "no limit"
template <class T> K7_HD T k7_curve_speed(const T& kap, double alpha)
{
    const double k = k7_val(kap), a = fabs(k);
    if (a < 1e-15) return k7_chain(kap, 1e5, 0.0);
    const double f = alpha / cbrt(a);
    return k7_chain(kap, f, -f / (3.0 * a) * (k < 0 ? -1.0 : 1.0));
}

// ------------------------------------------------------------------------------------------------------------
// problem data of one scenario
// ------------------------------------------------------------------------------------------------------------
struct K7Ctx {
    const eepacc_params* p;
    const eepacc_route* r;
    const eepacc_nlp_tables* tb;
    const double* s_tv;          // [N + 1]
    double m;                    // vehicle mass of this scenario
    double sc;                   // objective scaling
    int N, n_tl, nrow, flat;     // flat: sum(slope) < 0.1 -> theta = 0 (RunOpt_NLP.m:378-382)
};

K7_HD void k7_ctx_init(K7Ctx& c, const eepacc_params* p, const eepacc_route* r, const eepacc_nlp_tables* tb,
                       const double* s_tv, int N, double sc)
{
    c.p = p; c.r = r; c.tb = tb; c.s_tv = s_tv; c.N = N; c.sc = sc;
    c.m = p->m * r->mass_scale;
    c.n_tl = r->n_tl;
    c.nrow = K7_ROWS_BASE + 2 * r->n_tl;
    double ss = 0.0;
    for (int i = 0; i < r->n_slope; ++i) ss += r->slope[i];
    c.flat = ss < 1e-1;
}

template <class T> K7_HD T k7_theta(const K7Ctx& c, const T& s)
{
    if (c.flat) { T z; k7_const(z, 0.0); return z; }
    return k7_lut(s, c.r->s_slope, c.r->slope, c.r->n_slope);
}

// GetMotorPower_FifthOrderSurface.m / _QuadraticFit.m as used at RunOpt_NLP.m:228-240
template <class T> K7_HD T k7_power(const K7Ctx& c, const T& x, const T& y)
{
    if (c.p->nlp_use_fifth) {
        const double* b = c.p->b_fifth;
        T x2 = x * x, x3 = x2 * x, x4 = x3 * x, x5 = x4 * x;
        T y2 = y * y, y3 = y2 * y, y4 = y3 * y, y5 = y4 * y;
        return b[1] * x + b[2] * y + b[3] * x2 + b[4] * (x * y) + b[5] * y2 + b[6] * x3 + b[7] * (x2 * y) +
               b[8] * (x * y2) + b[9] * y3 + b[10] * x4 + b[11] * (x3 * y) + b[12] * (x2 * y2) + b[13] * (x * y3) +
               b[14] * y4 + b[15] * x5 + b[16] * (x4 * y) + b[17] * (x3 * y2) + b[18] * (x2 * y3) + b[19] * (x * y4) +
               b[20] * y5 + b[0];
    }
    const double* b = c.p->b_quadr;
    return b[1] * x + b[2] * y + b[3] * (x * x) + b[4] * (x * y) + b[5] * (y * y) + b[0];
}

// second partials of the power map w.r.t. (Fm, rpm), for the Hessian
K7_HD void k7_power_hess(const K7Ctx& c, double x, double y, double& pxx, double& pxy, double& pyy)
{
    if (c.p->nlp_use_fifth) {
        const double* b = c.p->b_fifth;
        const double x2 = x * x, x3 = x2 * x, y2 = y * y, y3 = y2 * y;
        pxx = 2 * b[3] + 6 * b[6] * x + 2 * b[7] * y + 12 * b[10] * x2 + 6 * b[11] * x * y + 2 * b[12] * y2 +
              20 * b[15] * x3 + 12 * b[16] * x2 * y + 6 * b[17] * x * y2 + 2 * b[18] * y3;
        pxy = b[4] + 2 * b[7] * x + 2 * b[8] * y + 3 * b[11] * x2 + 4 * b[12] * x * y + 3 * b[13] * y2 +
              4 * b[16] * x3 + 6 * b[17] * x2 * y + 6 * b[18] * x * y2 + 4 * b[19] * y3;
        pyy = 2 * b[5] + 2 * b[8] * x + 6 * b[9] * y + 2 * b[12] * x2 + 6 * b[13] * x * y + 12 * b[14] * y2 +
              2 * b[17] * x3 + 6 * b[18] * x2 * y + 12 * b[19] * x * y2 + 20 * b[20] * y3;
    } else {
        const double* b = c.p->b_quadr;
        pxx = 2 * b[3]; pxy = b[4]; pyy = 2 * b[5];
    }
}

// ------------------------------------------------------------------------------------------------------------
// the stage: interval k maps (x_k, u_k) to x_{k+1}, the interval's cost, and the rows d_i <= 0
// ------------------------------------------------------------------------------------------------------------
template <class T> struct K7Stage { T xn[K7_NX]; T cost; T row[K7_MAXR]; };

// f of RunOpt_NLP.m:225-245: (s, v)' and the integrand L, theta and j constant over the interval (xdot = 0 for both)
template <class T> K7_HD void k7_f(const K7Ctx& c, const T& v, const T& cth, const T& sth, const T& j, const T* u,
                                   T& ds, T& dv, T& L)
{
    const eepacc_params& p = *c.p;
    const double* W = p.W_NLP;
    const double m = c.m;
    T acc = (u[0] + u[1] - p.zeta_a * (v * v) - ((p.c_r * m * p.g) * cth + (m * p.g) * sth)) * (1.0 / (p.lambda * m));
    ds = v; dv = acc;
    T rpm = ((30.0 / K7_PI) * p.phi) * v;
    T P = k7_power(c, u[0], rpm);
    L = W[0] * P + W[1] * (acc * acc) + W[2] * (j * j) + W[3] * u[2] + W[4] * (u[3] * u[3] + 1e2 * u[3]) +
        W[5] * u[4] + W[6] * u[5];
}

// rows whose slack is xi_f / xi_s / xi_v / xi_h (used by the start heuristic); -1: no slack
K7_HD int k7_row_slack(int i)
{
    if (i >= 11 && i <= 22) return 5;
    if (i == 23 || i == 26 || i >= K7_ROWS_BASE) return 4;
    if (i == 24) return 2;
    if (i == 27) return 3;
    return -1;
}

template <class T, bool ROWS> K7_HD void k7_stage(const K7Ctx& c, int k, const T* x, const T* u, K7Stage<T>& o)
{
    const eepacc_params& p = *c.p;
    const double m = c.m;
    // ---- F: M = 4 RK4 steps (RunOpt_NLP.m:275-288); theta_k = slopeLookup(s_k) is a state that does not move
    T th = k7_theta(c, x[0]);
    const T cth = k7_cos(th), sth = k7_sin(th);
    T s = x[0], v = x[1];
    T Q; k7_const(Q, 0.0);
    const double DT = p.Ts / 4;
    for (int ii = 0; ii < 4; ++ii) {
        T k1s, k1v, k2s, k2v, k3s, k3v, k4s, k4v, q1, q2, q3, q4;
        k7_f(c, v, cth, sth, x[3], u, k1s, k1v, q1);
        k7_f(c, v + (DT / 2) * k1v, cth, sth, x[3], u, k2s, k2v, q2);
        k7_f(c, v + (DT / 2) * k2v, cth, sth, x[3], u, k3s, k3v, q3);
        k7_f(c, v + DT * k3v, cth, sth, x[3], u, k4s, k4v, q4);
        s = s + (DT / 6) * (k1s + 2.0 * k2s + 2.0 * k3s + k4s);
        v = v + (DT / 6) * (k1v + 2.0 * k2v + 2.0 * k3v + k4v);
        Q = Q + (DT / 6) * (q1 + 2.0 * q2 + 2.0 * q3 + q4);
    }
    o.cost = Q;
    // ---- x_{k+1}: theta+ = slopeLookup(s+) (:377-382), c+ and the jerk state (:383-386)
    T thn = k7_theta(c, s);
    T cn = k7_cos(thn), sn = k7_sin(thn);
    T cplus = u[0] + u[1] - p.zeta_a * (v * v) - (p.c_r * m * p.g) * cn - (m * p.g) * sn;
    o.xn[0] = s; o.xn[1] = v; o.xn[2] = cplus;
    o.xn[3] = (cplus - x[2]) * (1.0 / (p.lambda * m * p.Ts));
    if (!ROWS) return;
    // ---- rows, all as d <= 0
    const T& Fm = u[0]; const T& Fb = u[1];
    const T& xiv = u[2]; const T& xih = u[3]; const T& xis = u[4]; const T& xif = u[5];
    const double Fm_min = -p.phi * p.T_m_max / p.eta_TF, Fm_max = p.phi * p.T_m_max * p.eta_TF;   // :195-196
    T* d = o.row;
    d[0] = Fm - Fm_max;                                                       // :318-319 ubz
    d[1] = Fm_min - Fm;                                                       // lbz
    d[2] = Fb;                                                                // Fb_max = 0 (Fb_min = -inf: no row)
    d[3] = -xiv; d[4] = -xih; d[5] = -xis; d[6] = -xif;                       // slacks >= 0
    d[7] = -s;                                                                // :364 s_min = 0
    d[8] = -v;                                                                // v_min = 0
    d[9] = v - p.v_max;                                                       // v_max
    if (p.s_goal < 1e19) d[10] = s - p.s_goal; else k7_const(d[10], -1.0);    // s_max = s_goal (inf by default)
    T Fv = Fm * v;
    d[11] = -(Fv + (p.P_m_max / p.eta_TF) + xif);                             // :389-391 motor power, lower
    d[12] = Fv - (p.P_m_max * p.eta_TF) - xif;                                // :393-395
    T fr = (p.mu * m * p.g) * cn;
    d[13] = -(Fm + Fb + fr + xif);                                            // :398-400 total friction
    d[14] = Fm + Fb - fr - xif;                                               // :402-404
    T rear = (p.h_g / m) * cplus + (p.h_g * p.zeta_a / m) * (v * v) + p.g * (p.L_f * cn + p.h_g * sn);
    T lf = (p.L / (p.mu * m)) * Fm;
    d[15] = -(lf + rear + xif);                                               // :407-411 rear-axle traction
    d[16] = lf - rear - xif;                                                  // :413-417
    const double gv[4] = { 0, 5, 20, 25 };                                    // :74-80 ISO lookups
    const double amin[4] = { -4, -4, -2, -2 }, amax[4] = { 5, 5, 3.5, 3.5 }, jmag[4] = { 5, 5, 2.5, 2.5 };
    T a = cplus * (1.0 / (p.lambda * m));
    d[17] = -(a - k7_lut(v, gv, amin, 4) + xif);                              // :420-423
    d[18] = a - k7_lut(v, gv, amax, 4) - xif;                                 // :425-428
    T jl = k7_lut(v, gv, jmag, 4);
    d[19] = -(o.xn[3] + jl + xif);                                            // :431-433
    d[20] = o.xn[3] - jl - xif;                                               // :435-437
    d[21] = v - k7_lut(s, c.r->s_speedLim, c.r->v_speedLim, c.r->n_sl) - xif;             // :440-442
    d[22] = v - k7_curve_speed(k7_lut(s, c.r->s_curv, c.r->curvature, c.r->n_curv), p.alpha_TTL) - xif;   // :445-447
    d[23] = v - k7_lut(s, c.tb->s_stop, c.tb->v_stop, c.tb->n_stop) - xis;                // :450-452
    d[24] = -(v - k7_lut(s, c.tb->s_velInc, c.tb->v_velInc, c.tb->n_vi) + xiv);           // :473-475
    const double stv = c.s_tv[k];                                             // s_tv(k+1), 1-based (:478-480)
    d[25] = s - (stv - p.h_min);
    d[26] = s + p.tau_min * v - xis - stv;
    const double A_hwp = 2, T_hwp = 2, G_hwp = -0.0246 * T_hwp + 0.010819;    // :483-485
    d[27] = s + T_hwp * v + G_hwp * (v * v) - xih - (stv - A_hwp);            // :486-488
    const double D = p.stopRefDist, vi = D * p.stopRefVelSlope;               // :92, :138-146
    for (int i = 0; i < c.n_tl; ++i) {
        const double* tl = c.r->TLLoc[i];
        const double xs[3] = { tl[0] - D, tl[0], tl[0] + D }, ys[3] = { vi, p.TLstopVel, vi };
        T vp = k7_lut(s, xs, ys, 3);
        const double tt = k * p.Ts - tl[1], per = tl[2] + tl[3];
        const double ph = tt - floor(tt / per) * per;                         // MATLAB mod (:147)
        const double state = ph < tl[2] ? 0.2 : 1e3;                          // red .2 / green 1e3 (:148-151)
        d[K7_ROWS_BASE + 2 * i] = v - vp - state - xis;                       // :461-463
        d[K7_ROWS_BASE + 2 * i + 1] = -(v + vp + (1e3 - 10) - state + xis);   // :466-468
    }
}

// ------------------------------------------------------------------------------------------------------------
// the solver
// ------------------------------------------------------------------------------------------------------------
struct K7Work { double* ws; long long stride; };
#define K7_W(w, k, f) ((w).ws[((long long)(k) * K7_STAGE + (f)) * (w).stride])

struct K7Result { int status, iters, backtracks; double f, kkt, infeas, compl_, mu, delta; };

K7_HD void k7_seed(const double* x, const double* u, K7Dual* X, K7Dual* U)
{
    for (int i = 0; i < K7_NX; ++i) { k7_const(X[i], x[i]); X[i].d[i] = 1.0; }
    for (int i = 0; i < K7_NU; ++i) { k7_const(U[i], u[i]); U[i].d[K7_NX + i] = 1.0; }
}

K7_HD void k7_x0(const K7Ctx& c, double* x)
{
    const eepacc_params& p = *c.p;
    x[0] = p.s_init; x[1] = p.v_init;                                         // :212 x_init, U_{-1} = 0 (:338)
    const double th = k7_val(k7_theta(c, x[0]));
    x[2] = -p.zeta_a * x[1] * x[1] - p.c_r * c.m * p.g * cos(th) - c.m * p.g * sin(th);
    x[3] = 0.0;
}

// slacks and multipliers for the controls/states in the workspace: t = max(-d, tmin), lambda = mu / t
K7_HD void k7_init_duals(const K7Ctx& c, K7Work& w, double mu)
{
    for (int k = 0; k < c.N; ++k) {
        double x[K7_NX], u[K7_NU];
        for (int i = 0; i < K7_NX; ++i) x[i] = K7_W(w, k, K7_F_X + i);
        for (int i = 0; i < K7_NU; ++i) u[i] = K7_W(w, k, K7_F_U + i);
        K7Stage<double> o;
        k7_stage<double, true>(c, k, x, u, o);
        for (int i = 0; i < K7_NX; ++i) K7_W(w, k + 1, K7_F_X + i) = o.xn[i];
        for (int i = 0; i < c.nrow; ++i) {
            const double t = fmax(-o.row[i], 1e-2);
            K7_W(w, k, K7_F_T + i) = t;
            K7_W(w, k, K7_F_LAM + i) = mu / t;
        }
    }
}

// built-in start (the reference starts IPOPT from z0 = 0, :322,:367; any start leads to the same KKT conditions):
// a damped car-following roll-out behind the lead vehicle, capped by the speed limits, slacks sized to their rows
K7_HD void k7_start(const K7Ctx& c, K7Work& w)
{
    const eepacc_params& p = *c.p;
    double x[K7_NX];
    k7_x0(c, x);
    const double Fm_min = -p.phi * p.T_m_max / p.eta_TF, Fm_max = p.phi * p.T_m_max * p.eta_TF;
    for (int i = 0; i < K7_NX; ++i) K7_W(w, 0, K7_F_X + i) = x[i];
    for (int k = 0; k < c.N; ++k) {
        const double s = x[0], v = x[1];
        const double stv = c.s_tv[k], stv1 = c.s_tv[k + 1 <= c.N ? k + 1 : c.N];
        const double vl = fmax((stv1 - stv) / p.Ts, 0.0);
        const double gap = stv - s - 3.0 - 2.2 * v;
        double a = 0.08 * gap + 0.6 * (vl - v);
        a = fmin(fmax(a, -2.5), 1.0);
        double vcap = k7_lut(s + 2.0 * v, c.r->s_speedLim, c.r->v_speedLim, c.r->n_sl);
        vcap = fmin(vcap, k7_curve_speed(k7_lut(s + 2.0 * v, c.r->s_curv, c.r->curvature, c.r->n_curv), p.alpha_TTL));
        vcap = fmin(vcap, k7_lut(s + v, c.tb->s_stop, c.tb->v_stop, c.tb->n_stop));
        vcap = 0.97 * fmin(vcap, p.v_max);
        if (v + a * p.Ts > vcap) a = fmax((vcap - v) / p.Ts, -3.0);
        if (v + a * p.Ts < 0.05) a = (0.05 - v) / p.Ts;
        const double th = k7_val(k7_theta(c, s));
        double u[K7_NU] = { 0, 0, 0, 0, 0, 0 };
        double Fm = p.lambda * c.m * a + p.zeta_a * v * v + c.m * p.g * (p.c_r * cos(th) + sin(th));
        Fm = fmin(fmax(Fm, 0.9 * Fm_min), 0.9 * Fm_max);
        if (Fm * fmax(v, 1.0) > 0.9 * p.P_m_max * p.eta_TF) Fm = 0.9 * p.P_m_max * p.eta_TF / fmax(v, 1.0);
        u[0] = Fm;
        K7Stage<double> o;
        k7_stage<double, true>(c, k, x, u, o);
        double need[K7_NU] = { 0, 0, 0, 0, 0, 0 };
        for (int i = 0; i < c.nrow; ++i) {
            const int sl = k7_row_slack(i);
            if (sl >= 0) need[sl] = fmax(need[sl], o.row[i]);
        }
        for (int i = 2; i < K7_NU; ++i) u[i] = need[i] + 0.1;
        for (int i = 0; i < K7_NU; ++i) K7_W(w, k, K7_F_U + i) = u[i];
        for (int i = 0; i < K7_NX; ++i) x[i] = o.xn[i];
    }
}

// 6 x 6 Cholesky in place (lower); false if a pivot is not safely positive
K7_HD bool k7_chol6(double (*A)[K7_NU])
{
    for (int j = 0; j < K7_NU; ++j) {
        double dj = A[j][j];
        for (int l = 0; l < j; ++l) dj -= A[j][l] * A[j][l];
        if (!(dj > 1e-14 * (1.0 + fabs(A[j][j])))) return false;
        dj = sqrt(dj);
        A[j][j] = dj;
        for (int i = j + 1; i < K7_NU; ++i) {
            double sij = A[i][j];
            for (int l = 0; l < j; ++l) sij -= A[i][l] * A[j][l];
            A[i][j] = sij / dj;
        }
    }
    return true;
}
K7_HD void k7_chol6_solve(const double (*Lc)[K7_NU], double* b)
{
    for (int i = 0; i < K7_NU; ++i) { double s = b[i]; for (int l = 0; l < i; ++l) s -= Lc[i][l] * b[l]; b[i] = s / Lc[i][i]; }
    for (int i = K7_NU - 1; i >= 0; --i) { double s = b[i]; for (int l = i + 1; l < K7_NU; ++l) s -= Lc[l][i] * b[l]; b[i] = s / Lc[i][i]; }
}

#ifdef K7_TRACE
static int g_k7_arg_k, g_k7_arg_a;
#endif
struct K7Errs { double stat, feas, comp_mu, comp_0, theta, sum_lam, sum_pi, fscaled, max_lam; };

// Backward sweep at the current iterate: residuals, gains.  Returns false when a stage's reduced Hessian is not
// positive definite under the shift `delta` (the caller raises it: inertia correction, Waechter & Biegler 3.1).
K7_HD bool k7_backward(const K7Ctx& c, K7Work& w, double mu, double delta, K7Errs& e)
{
    const eepacc_params& p = *c.p;
    const double* W = p.W_NLP;
    const double lm = p.lambda * c.m, a_vv = -2.0 * p.zeta_a / lm;
    const double c1 = (30.0 / K7_PI) * p.phi;
    const double G_hwp = -0.0246 * 2 + 0.010819;
    double P[K7_NX][K7_NX], pv[K7_NX], pi[K7_NX];
    for (int i = 0; i < K7_NX; ++i) { pv[i] = 0.0; pi[i] = 0.0; for (int j = 0; j < K7_NX; ++j) P[i][j] = 0.0; }
    e.stat = e.feas = e.comp_mu = e.comp_0 = e.theta = e.sum_lam = e.sum_pi = e.fscaled = e.max_lam = 0.0;
    for (int k = c.N - 1; k >= 0; --k) {
        double x[K7_NX], u[K7_NU];
        for (int i = 0; i < K7_NX; ++i) x[i] = K7_W(w, k, K7_F_X + i);
        for (int i = 0; i < K7_NU; ++i) u[i] = K7_W(w, k, K7_F_U + i);
        K7Dual X[K7_NX], U[K7_NU];
        k7_seed(x, u, X, U);
        K7Stage<K7Dual> o;
        k7_stage<K7Dual, true>(c, k, X, U, o);
        double Q[K7_NZ][K7_NZ], q[K7_NZ], rz[K7_NZ];
        for (int a = 0; a < K7_NZ; ++a) {
            const double g = c.sc * o.cost.d[a];
            K7_W(w, k, K7_F_GL + a) = g;
            double qa = g, ra = g;
            for (int i = 0; i < K7_NX; ++i) { qa += o.xn[i].d[a] * pv[i]; ra += o.xn[i].d[a] * pi[i]; }
            q[a] = qa; rz[a] = ra;
        }
        K7_W(w, k, K7_F_L0) = c.sc * o.cost.v;
        e.fscaled += c.sc * o.cost.v;
        // dPhi' P+ dPhi
        double PA[K7_NX][K7_NZ];
        for (int i = 0; i < K7_NX; ++i)
            for (int b = 0; b < K7_NZ; ++b) {
                double s = 0.0;
                for (int j = 0; j < K7_NX; ++j) s += P[i][j] * o.xn[j].d[b];
                PA[i][b] = s;
            }
        for (int a = 0; a < K7_NZ; ++a)
            for (int b = 0; b <= a; ++b) {
                double s = 0.0;
                for (int i = 0; i < K7_NX; ++i) s += o.xn[i].d[a] * PA[i][b];
                Q[a][b] = s;
            }
        // Hessian of the Lagrangian (lower triangle), z = (s, v, c, j, Fm, Fb, xi_v, xi_h, xi_s, xi_f)
        {
            const double v = x[1], Fm = u[0];
            double pxx, pxy, pyy;
            k7_power_hess(c, Fm, c1 * v, pxx, pxy, pyy);
            const double th = k7_val(k7_theta(c, x[0]));
            const double acc = (u[0] + u[1] - p.zeta_a * v * v - c.m * p.g * (p.c_r * cos(th) + sin(th))) / lm;
            const double av = -2.0 * p.zeta_a * v / lm, af = 1.0 / lm;
            const double h = c.sc * p.Ts;
            Q[1][1] += h * (W[0] * pyy * c1 * c1 + 2.0 * W[1] * (av * av + acc * a_vv));
            Q[4][1] += h * (W[0] * pxy * c1 + 2.0 * W[1] * av * af);
            Q[5][1] += h * (2.0 * W[1] * av * af);
            Q[4][4] += h * (W[0] * pxx + 2.0 * W[1] * af * af);
            Q[5][4] += h * (2.0 * W[1] * af * af);
            Q[5][5] += h * (2.0 * W[1] * af * af);
            Q[3][3] += h * 2.0 * W[2];
            Q[7][7] += h * 2.0 * W[4];
            // curvature of the dynamics weighted by the costates: v+ ~ v + Ts a, s+ ~ s + Ts v + Ts^2 a / 2
            Q[1][1] += (pi[1] * p.Ts + pi[0] * 0.5 * p.Ts * p.Ts) * a_vv;
        }
        const double* lam_base = &K7_W(w, k, K7_F_LAM);
        const double* t_base = &K7_W(w, k, K7_F_T);
        double wvv = 0.0, wfv = 0.0;           // coefficients of dv+ dv+' and of (e_Fm dv+' + dv+ e_Fm')
        wvv += (pi[2] + pi[3] / (lm * p.Ts)) * (-2.0 * p.zeta_a);
        for (int i = 0; i < c.nrow; ++i) {
            const double lam = lam_base[(long long)i * w.stride], t = t_base[(long long)i * w.stride];
            const double d = o.row[i].v, rp = d + t;
            K7_W(w, k, K7_F_RP + i) = rp;
            const double sig = lam / t, lh = (mu + lam * rp) / t;
            const double* J = o.row[i].d;
            for (int a = 0; a < K7_NZ; ++a) {
                const double sa = sig * J[a];
                q[a] += lh * J[a];
                rz[a] += lam * J[a];
                if (sa != 0.0) for (int b = 0; b <= a; ++b) Q[a][b] += sa * J[b];
            }
            e.feas = fmax(e.feas, fabs(rp));
            e.theta += fabs(rp);
            e.comp_mu = fmax(e.comp_mu, fabs(lam * t - mu));
            e.comp_0 = fmax(e.comp_0, fabs(lam * t));
            e.sum_lam += fabs(lam);
            e.max_lam = fmax(e.max_lam, lam);
            if (i == 27) wvv += lam * 2.0 * G_hwp;
            if (i == 11) wfv -= lam;
            if (i == 12) wfv += lam;
            if (i == 17) wvv += lam * 2.0 * p.zeta_a / lm;
            if (i == 18) wvv -= lam * 2.0 * p.zeta_a / lm;
        }
        {
            const double* dv = o.xn[1].d;
            for (int a = 0; a < K7_NZ; ++a)
                for (int b = 0; b <= a; ++b) {
                    double s = wvv * dv[a] * dv[b];
                    if (a == 4) s += wfv * dv[b];
                    if (b == 4) s += wfv * dv[a];
                    Q[a][b] += s;
                }
        }
        for (int a = 0; a < K7_NZ; ++a) Q[a][a] += delta;
        // ---- Riccati step: factor Quu, gains, value function of stage k
        double Lc[K7_NU][K7_NU];
        for (int a = 0; a < K7_NU; ++a)
            for (int b = 0; b < K7_NU; ++b) Lc[a][b] = a >= b ? Q[K7_NX + a][K7_NX + b] : Q[K7_NX + b][K7_NX + a];
        if (!k7_chol6(Lc)) {
#ifdef K7_TRACE
            printf("   chol fails at stage %d delta %.2e: Quu diag %.3e %.3e %.3e %.3e %.3e %.3e  P diag %.3e %.3e %.3e %.3e  W11 %.3e\n", k, delta,
                   Q[4][4], Q[5][5], Q[6][6], Q[7][7], Q[8][8], Q[9][9], P[0][0], P[1][1], P[2][2], P[3][3], Q[1][1]);
#endif
            return false;
        }
        double kff[K7_NU], Kfb[K7_NX][K7_NU];      // Kfb[i] = column i of K (6 entries)
        for (int a = 0; a < K7_NU; ++a) kff[a] = -q[K7_NX + a];
        k7_chol6_solve(Lc, kff);
        for (int i = 0; i < K7_NX; ++i) {
            for (int a = 0; a < K7_NU; ++a) Kfb[i][a] = -Q[K7_NX + a][i];
            k7_chol6_solve(Lc, Kfb[i]);
        }
        for (int a = 0; a < K7_NU; ++a) {
            K7_W(w, k, K7_F_KFF + a) = kff[a];
            for (int i = 0; i < K7_NX; ++i) K7_W(w, k, K7_F_KFB + a * K7_NX + i) = Kfb[i][a];
        }
        double Pn[K7_NX][K7_NX], pn[K7_NX];
        for (int i = 0; i < K7_NX; ++i) {
            double s = q[i];
            for (int a = 0; a < K7_NU; ++a) s += Q[K7_NX + a][i] * kff[a];
            pn[i] = s;
            for (int j = 0; j <= i; ++j) {
                double sij = Q[i][j];
                for (int a = 0; a < K7_NU; ++a) sij += Q[K7_NX + a][i] * Kfb[j][a];
                Pn[i][j] = sij;
            }
        }
        for (int i = 0; i < K7_NX; ++i) {
            pv[i] = pn[i];
            for (int j = 0; j <= i; ++j) { P[i][j] = Pn[i][j]; P[j][i] = Pn[i][j]; }
        }
#ifdef K7_TRACE
        for (int a = 0; a < K7_NU; ++a) if (fabs(rz[K7_NX + a]) > e.stat) { g_k7_arg_k = k; g_k7_arg_a = a; }
#endif
        for (int a = 0; a < K7_NU; ++a) e.stat = fmax(e.stat, fabs(rz[K7_NX + a]));
        for (int i = 0; i < K7_NX; ++i) { pi[i] = rz[i]; e.sum_pi += fabs(rz[i]); }
    }
    return true;
}

// Trial roll-out with step alpha; fills the trial fields.  The slack of a row follows t+ = (1 - alpha)(d + t) - d(z+), which
// shrinks the row's residual by exactly (1 - alpha).  Returns false if a slack would fall below the fraction-to-the-boundary
// floor (1 - tau) t.  dphi: change of f - mu sum log t; dlin: its first-order model at the actual displacement.
K7_HD bool k7_trial(const K7Ctx& c, K7Work& w, double mu, double alpha, double tau, double& dphi, double& dlin)
{
    double xt[K7_NX];
    for (int i = 0; i < K7_NX; ++i) { xt[i] = K7_W(w, 0, K7_F_X + i); K7_W(w, 0, K7_F_XT + i) = xt[i]; }
    dphi = 0.0; dlin = 0.0;
    for (int k = 0; k < c.N; ++k) {
        double dx[K7_NX], ut[K7_NU];
        for (int i = 0; i < K7_NX; ++i) dx[i] = xt[i] - K7_W(w, k, K7_F_X + i);
        double lin = 0.0;
        for (int i = 0; i < K7_NX; ++i) lin += K7_W(w, k, K7_F_GL + i) * dx[i];
        for (int a = 0; a < K7_NU; ++a) {
            double du = alpha * K7_W(w, k, K7_F_KFF + a);
            for (int i = 0; i < K7_NX; ++i) du += K7_W(w, k, K7_F_KFB + a * K7_NX + i) * dx[i];
            ut[a] = K7_W(w, k, K7_F_U + a) + du;
            K7_W(w, k, K7_F_UT + a) = ut[a];
            lin += K7_W(w, k, K7_F_GL + K7_NX + a) * du;
        }
        K7Stage<double> o;
        k7_stage<double, true>(c, k, xt, ut, o);
        double bar = 0.0, blin = 0.0;
        for (int i = 0; i < c.nrow; ++i) {
            const double t = K7_W(w, k, K7_F_T + i), rp = K7_W(w, k, K7_F_RP + i);
            const double tn = (1.0 - alpha) * rp - o.row[i];
            if (!(tn >= (1.0 - tau) * t)) return false;
            K7_W(w, k, K7_F_TT + i) = tn;
            bar += log(tn / t);
            blin += (tn - t) / t;
        }
        dphi += (c.sc * o.cost - K7_W(w, k, K7_F_L0)) - mu * bar;
        dlin += lin - mu * blin;
        for (int i = 0; i < K7_NX; ++i) { xt[i] = o.xn[i]; K7_W(w, k + 1, K7_F_XT + i) = xt[i]; }
    }
    return true;
}

// Accept the trial point: multiplier step with its own fraction-to-the-boundary length, then copy.
K7_HD void k7_accept(const K7Ctx& c, K7Work& w, double mu, double alpha, double tau)
{
    double ad = 1.0;
    for (int k = 0; k < c.N; ++k)
        for (int i = 0; i < c.nrow; ++i) {
            const double t = K7_W(w, k, K7_F_T + i), tn = K7_W(w, k, K7_F_TT + i), lam = K7_W(w, k, K7_F_LAM + i);
            const double dl = (mu - lam * t - lam * (tn - t) / alpha) / t;
            if (dl < 0.0) ad = fmin(ad, -tau * lam / dl);
        }
    for (int k = 0; k < c.N; ++k) {
        for (int i = 0; i < c.nrow; ++i) {
            const double t = K7_W(w, k, K7_F_T + i), tn = K7_W(w, k, K7_F_TT + i), lam = K7_W(w, k, K7_F_LAM + i);
            const double dl = (mu - lam * t - lam * (tn - t) / alpha) / t;
            double ln = lam + ad * dl;
            ln = fmin(fmax(ln, mu / (1e10 * tn)), 1e10 * mu / tn);            // Waechter & Biegler eq. (16)
            K7_W(w, k, K7_F_LAM + i) = ln;
            K7_W(w, k, K7_F_T + i) = tn;
        }
        for (int a = 0; a < K7_NU; ++a) K7_W(w, k, K7_F_U + a) = K7_W(w, k, K7_F_UT + a);
        for (int i = 0; i < K7_NX; ++i) K7_W(w, k + 1, K7_F_X + i) = K7_W(w, k + 1, K7_F_XT + i);
    }
}

// The interior-point loop.  The workspace must hold x_0 and u_0..u_{N-1} (k7_start or the caller's guess).
K7_HD void k7_solve(const K7Ctx& c, K7Work& w, const eepacc_nlp_opts& op, K7Result& res)
{
    double mu = op.mu0, delta = 0.0, nu = 1.0;
    k7_init_duals(c, w, mu);
    res.status = EEPACC_NLP_MAXITER; res.iters = 0; res.backtracks = 0;
    K7Errs e;
    int fails = 0, stall = 0, tiny = 0;
    double delta_last = 0.0;
    for (int it = 0; it < op.max_iter; ++it) {
        res.iters = it;
        // ---- Newton system at the current iterate, with inertia correction
        int tries = 0;
        while (!k7_backward(c, w, mu, delta, e)) {
            delta = delta == 0.0 ? fmax(1e-8, delta_last / 3.0) : delta * 8.0;
            if (++tries > 40) { res.status = EEPACC_NLP_NUMERIC; goto done; }
        }
        if (!(e.stat == e.stat) || !(e.feas == e.feas)) { res.status = EEPACC_NLP_NUMERIC; goto done; }
        {
            const double nmul = (double)c.N * (c.nrow + K7_NX);
            const double sd = fmax(100.0, (e.sum_lam + e.sum_pi) / nmul) / 100.0;
            const double scs = fmax(100.0, e.sum_lam / ((double)c.N * c.nrow)) / 100.0;
            const double err0 = fmax(fmax(e.stat / sd, e.feas), e.comp_0 / scs);
            const double errm = fmax(fmax(e.stat / sd, e.feas), e.comp_mu / scs);
            res.kkt = err0; res.infeas = e.feas; res.compl_ = e.comp_0 / scs; res.mu = mu; res.delta = delta;
            res.f = e.fscaled / c.sc;
#ifdef K7_TRACE
            printf("it %3d mu %.2e stat %.3e feas %.3e comp %.3e theta %.3e f %.10e delta %.1e nu %.2e sd %.2e\n", it, mu, e.stat / sd,
                   e.feas, e.comp_0 / scs, e.theta, res.f, delta, nu, sd);
            printf("       worst stationarity at stage %d control %d: u = %.4e %.4e %.4e %.4e %.4e %.4e\n", g_k7_arg_k, g_k7_arg_a,
                   K7_W(w, g_k7_arg_k, K7_F_U), K7_W(w, g_k7_arg_k, K7_F_U + 1), K7_W(w, g_k7_arg_k, K7_F_U + 2), K7_W(w, g_k7_arg_k, K7_F_U + 3),
                   K7_W(w, g_k7_arg_k, K7_F_U + 4), K7_W(w, g_k7_arg_k, K7_F_U + 5));
#endif
            if (err0 <= op.tol) { res.status = EEPACC_NLP_OK; goto done; }
            // The barrier problem counts as solved at its tolerance (Waechter & Biegler eq. (7)) or when three accepted
            // steps in a row moved the merit function by less than its round-off: the slack of a tight headway row is
            // recomputed from positions of ~3e3 m, whose last bit (4.5e-13 m) times lambda / t is the floor of the
            // stationarity residual.  At the last mu a stalled iteration ends the solve as "acceptable", the counterpart
            // of the reference's opts.ipopt.acceptable_tol / acceptable_obj_change_tol (RunOpt_NLP.m:253-254).
            if (errm <= 10.0 * mu || (stall >= 3 && err0 <= op.acceptable_kkt)) {
                if (mu > op.mu_min) {
                    mu = fmax(op.mu_min, fmin(K7_KAPPA_MU * mu, pow(mu, K7_THETA_MU)));
                    stall = 0;
                    continue;
                }
                if (stall >= 3) { res.status = EEPACC_NLP_ACCEPTABLE; goto done; }
            }
        }
        // ---- line search on phi = f - mu sum log t + nu |d + t|_1 along the closed-loop roll-out
        {
            const double tau = fmax(0.99, 1.0 - mu);
            double alpha = 1.0, dphi = 0.0, dlin = 0.0;
            bool ok = false;
            for (int ls = 0; ls < 50; ++ls, alpha *= 0.5) {
                if (ls) ++res.backtracks;
                if (!k7_trial(c, w, mu, alpha, tau, dphi, dlin)) continue;
                if (e.theta > 1e-12) {
                    const double need = dlin / (0.9 * alpha * e.theta);
                    if (nu < need) nu = need + 1.0;
                }
                const double pred = dlin - nu * alpha * e.theta;
                const double slack = 1e-13 * (1.0 + fabs(e.fscaled));
                if (dphi - nu * alpha * e.theta <= 1e-4 * fmin(pred, 0.0) + slack) { ok = true; break; }
            }
            if (!ok) {
                // no acceptable step along this direction: stiffen the model (larger shift) and try again
                delta = delta == 0.0 ? 1e-6 : delta * 10.0;
                if (++fails > 12) { res.status = EEPACC_NLP_LINESEARCH; goto done; }
                continue;
            }
            fails = 0;
            // thirty accepted steps in a row shorter than 1e-5: the iterate is jammed against a corner of a lookup table
            //  and would only burn the iteration limit -- and hold up its warp
            tiny = alpha < 1e-5 ? tiny + 1 : 0;
            if (tiny >= 30) { k7_accept(c, w, mu, alpha, tau); res.status = EEPACC_NLP_LINESEARCH; goto done; }
            stall = fabs(dphi) <= 1e-12 * (1.0 + fabs(e.fscaled)) ? stall + 1 : 0;
#ifdef K7_TRACE
            printf("       alpha %.3e dphi %.3e dlin %.3e\n", alpha, dphi, dlin);
#endif
            k7_accept(c, w, mu, alpha, tau);
            delta_last = delta > 0.0 ? delta : delta_last;
            delta = 0.0;
        }
    }
done:
    return;
}

// z in the reference's layout (RunOpt_NLP.m:533-543) from the workspace
K7_HD void k7_write_z(const K7Ctx& c, const K7Work& w, double* z)
{
    const double lmT = c.p->lambda * c.m * c.p->Ts;
    (void)lmT;
    for (int k = 0; k <= c.N; ++k) {
        double* X = z + (k == 0 ? 0 : 4 + 10 * (k - 1) + 6);
        const double s = K7_W(w, k, K7_F_X + 0);
        X[0] = s; X[1] = K7_W(w, k, K7_F_X + 1);
        X[2] = k7_val(k7_theta(c, s));
        X[3] = K7_W(w, k, K7_F_X + 3);
        if (k < c.N) {
            double* U = z + 4 + 10 * k;
            for (int a = 0; a < K7_NU; ++a) U[a] = K7_W(w, k, K7_F_U + a);
        }
    }
}

// controls of a caller-supplied guess into the workspace (states are re-generated by the roll-out)
K7_HD void k7_read_guess(const K7Ctx& c, K7Work& w, const double* z)
{
    double x[K7_NX];
    k7_x0(c, x);
    for (int i = 0; i < K7_NX; ++i) K7_W(w, 0, K7_F_X + i) = x[i];
    for (int k = 0; k < c.N; ++k)
        for (int a = 0; a < K7_NU; ++a) K7_W(w, k, K7_F_U + a) = z[4 + 10 * k + a];
}
