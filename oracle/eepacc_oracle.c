/*
 * eepacc_oracle.c -- CPU ORACLE (test infrastructure, not a product path; see eepacc_oracle.h).
 *
 * Restates, in plain C, the reference's MATLAB hot path.  Citation shorthand:
 *   MC/  = /root/reference/MATLAB_CasADi/        ABO/ = /root/reference/ACCMPC-ABO_CasADi/
 *   F/   = <tree>/Functions/
 */
#include "eepacc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define N ORC_N
static const double PI_ = 3.14159265358979323846;

/* ------------------------------------------------------------------------------------------
 * F/PWA_function_manipulation/InterpPWA.m:14-27
 * ------------------------------------------------------------------------------------------ */
double orc_interp_pwa(double d, const double* doms, const double* vals, int n)
{
    if (n <= 0) return 0.0;
    if (d < doms[0]) return vals[0];
    if (d > doms[n - 1]) return vals[n - 1];
    for (int i = 0; i < n - 1; ++i) {
        if (d >= doms[i] && d <= doms[i + 1]) {
            double f = (d - doms[i]) / (doms[i + 1] - doms[i]);
            return vals[i] + f * (vals[i + 1] - vals[i]);
        }
    }
    return vals[n - 1];   /* n == 1 and d == doms[0] */
}

/* ------------------------------------------------------------------------------------------
 * F/MPCs/RunPlantModel.m:20-44 -- M RK4 sub-steps of the nonlinear longitudinal dynamics
 * ------------------------------------------------------------------------------------------ */
static void plant_rhs(const eepacc_params* p, const eepacc_route* r, double m, const double x[2],
                      double u, double xd[2])
{
    double theta = orc_interp_pwa(x[0], r->s_slope, r->slope, r->n_slope);
    xd[0] = x[1];
    xd[1] = 1.0 / p->lambda / m * (u - p->zeta_a * x[1] * x[1] - p->c_r * m * p->g * cos(theta)
                                   - m * p->g * sin(theta));
}

void orc_plant(const eepacc_params* p, const eepacc_route* r, const double x0[2], const double u[2],
               double xn[2])
{
    const int M = p->n_integrate_plant;
    const double m = p->m * r->mass_scale;
    const double uk = u[0] + u[1];
    const double DT = p->Ts / M;
    double x[2] = { x0[0], x0[1] };
    for (int k = 0; k < M; ++k) {
        double k1[2], k2[2], k3[2], k4[2], t[2];
        plant_rhs(p, r, m, x, uk, k1);
        t[0] = x[0] + DT / 2 * k1[0]; t[1] = x[1] + DT / 2 * k1[1];
        plant_rhs(p, r, m, t, uk, k2);
        t[0] = x[0] + DT / 2 * k2[0]; t[1] = x[1] + DT / 2 * k2[1];
        plant_rhs(p, r, m, t, uk, k3);
        t[0] = x[0] + DT * k3[0]; t[1] = x[1] + DT * k3[1];
        plant_rhs(p, r, m, t, uk, k4);
        x[0] = x[0] + DT / 6 * (k1[0] + 2 * k2[0] + 2 * k3[0] + k4[0]);
        x[1] = x[1] + DT / 6 * (k1[1] + 2 * k2[1] + 2 * k3[1] + k4[1]);
    }
    xn[0] = x[0]; xn[1] = x[1];
}

/* ------------------------------------------------------------------------------------------
 * F/Other/GetMotorPower_FifthOrderSurface.m:16-20 and GetMotorPower_QuadraticFit.m:13
 * ------------------------------------------------------------------------------------------ */
double orc_power5(const double b[21], double x, double y)
{
    double x2 = x * x, x3 = x2 * x, x4 = x3 * x, x5 = x4 * x;
    double y2 = y * y, y3 = y2 * y, y4 = y3 * y, y5 = y4 * y;
    return b[0] + b[1] * x + b[2] * y + b[3] * x2 + b[4] * x * y + b[5] * y2 +
           b[6] * x3 + b[7] * x2 * y + b[8] * x * y2 + b[9] * y3 + b[10] * x4 +
           b[11] * x3 * y + b[12] * x2 * y2 + b[13] * x * y3 + b[14] * y4 +
           b[15] * x5 + b[16] * x4 * y + b[17] * x3 * y2 + b[18] * x2 * y3 +
           b[19] * x * y4 + b[20] * y5;
}

double orc_power2(const double b[6], double Fm, double rpm)
{
    return b[0] + b[1] * Fm + b[2] * rpm + b[3] * Fm * Fm + b[4] * Fm * rpm + b[5] * rpm * rpm;
}

/* ------------------------------------------------------------------------------------------
 * MC/RunOpt_NLP.m:225-245,263 (f) and :275-288 (F, M=4 RK4 steps); forward-mode Jacobian.
 * A tiny dual-number type (value + ND directional derivatives) gives f, F and dF in one code.
 * ------------------------------------------------------------------------------------------ */
#define ND 10
typedef struct { double v; double d[ND]; } dual;
static dual dconst(double c) { dual r; r.v = c; memset(r.d, 0, sizeof r.d); return r; }
static dual dadd(dual a, dual b) { for (int i = 0; i < ND; ++i) a.d[i] += b.d[i]; a.v += b.v; return a; }
static dual dsub(dual a, dual b) { for (int i = 0; i < ND; ++i) a.d[i] -= b.d[i]; a.v -= b.v; return a; }
static dual dmul(dual a, dual b)
{
    dual r; r.v = a.v * b.v;
    for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
    return r;
}
static dual dscale(double c, dual a) { for (int i = 0; i < ND; ++i) a.d[i] *= c; a.v *= c; return a; }
static dual dcos(dual a) { dual r; r.v = cos(a.v); double s = -sin(a.v); for (int i = 0; i < ND; ++i) r.d[i] = s * a.d[i]; return r; }
static dual dsin(dual a) { dual r; r.v = sin(a.v); double c = cos(a.v); for (int i = 0; i < ND; ++i) r.d[i] = c * a.d[i]; return r; }

static void f_dual(const eepacc_params* p, double m, const dual x[4], const dual u[6], dual xdot[4], dual* L)
{
    const double* W = p->W_NLP;
    dual v = x[1], th = x[2], j = x[3], Fm = u[0], Fb = u[1];
    /* xdot (RunOpt_NLP.m:225) */
    dual acc = dadd(Fm, Fb);
    acc = dsub(acc, dscale(p->zeta_a, dmul(v, v)));
    acc = dsub(acc, dscale(p->c_r * m * p->g, dcos(th)));
    acc = dsub(acc, dscale(m * p->g, dsin(th)));
    acc = dscale(1.0 / (p->lambda * m), acc);
    xdot[0] = v; xdot[1] = acc; xdot[2] = dconst(0.0); xdot[3] = dconst(0.0);
    /* P_bat (RunOpt_NLP.m:228-240) */
    dual rpm = dscale((30.0 / PI_) * p->phi, v);
    dual P;
    if (p->nlp_use_fifth) {
        const double* b = p->b_fifth;
        dual x1 = Fm, y1 = rpm;
        dual x2 = dmul(x1, x1), x3 = dmul(x2, x1), x4 = dmul(x3, x1), x5 = dmul(x4, x1);
        dual y2 = dmul(y1, y1), y3 = dmul(y2, y1), y4 = dmul(y3, y1), y5 = dmul(y4, y1);
        P = dconst(b[0]);
        P = dadd(P, dscale(b[1], x1));  P = dadd(P, dscale(b[2], y1));
        P = dadd(P, dscale(b[3], x2));  P = dadd(P, dscale(b[4], dmul(x1, y1)));
        P = dadd(P, dscale(b[5], y2));  P = dadd(P, dscale(b[6], x3));
        P = dadd(P, dscale(b[7], dmul(x2, y1)));  P = dadd(P, dscale(b[8], dmul(x1, y2)));
        P = dadd(P, dscale(b[9], y3));  P = dadd(P, dscale(b[10], x4));
        P = dadd(P, dscale(b[11], dmul(x3, y1))); P = dadd(P, dscale(b[12], dmul(x2, y2)));
        P = dadd(P, dscale(b[13], dmul(x1, y3))); P = dadd(P, dscale(b[14], y4));
        P = dadd(P, dscale(b[15], x5)); P = dadd(P, dscale(b[16], dmul(x4, y1)));
        P = dadd(P, dscale(b[17], dmul(x3, y2))); P = dadd(P, dscale(b[18], dmul(x2, y3)));
        P = dadd(P, dscale(b[19], dmul(x1, y4))); P = dadd(P, dscale(b[20], y5));
    } else {
        const double* b = p->b_quadr;
        P = dconst(b[0]);
        P = dadd(P, dscale(b[1], Fm)); P = dadd(P, dscale(b[2], rpm));
        P = dadd(P, dscale(b[3], dmul(Fm, Fm))); P = dadd(P, dscale(b[4], dmul(Fm, rpm)));
        P = dadd(P, dscale(b[5], dmul(rpm, rpm)));
    }
    /* L (RunOpt_NLP.m:245) */
    dual Lv = dscale(W[0], P);
    Lv = dadd(Lv, dscale(W[1], dmul(acc, acc)));
    Lv = dadd(Lv, dscale(W[2], dmul(j, j)));
    Lv = dadd(Lv, dscale(W[3], u[2]));
    Lv = dadd(Lv, dscale(W[4], dadd(dmul(u[3], u[3]), dscale(1e2, u[3]))));
    Lv = dadd(Lv, dscale(W[5], u[4]));
    Lv = dadd(Lv, dscale(W[6], u[5]));
    *L = Lv;
}

static void F_dual(const eepacc_params* p, double m, dual X[4], const dual U[6], dual* Q)
{
    const int M = 4;
    const double DT = p->Ts / M;
    *Q = dconst(0.0);
    for (int ii = 0; ii < M; ++ii) {
        dual k1[4], k2[4], k3[4], k4[4], t[4], q1, q2, q3, q4;
        f_dual(p, m, X, U, k1, &q1);
        for (int i = 0; i < 4; ++i) t[i] = dadd(X[i], dscale(DT / 2, k1[i]));
        f_dual(p, m, t, U, k2, &q2);
        for (int i = 0; i < 4; ++i) t[i] = dadd(X[i], dscale(DT / 2, k2[i]));
        f_dual(p, m, t, U, k3, &q3);
        for (int i = 0; i < 4; ++i) t[i] = dadd(X[i], dscale(DT, k3[i]));
        f_dual(p, m, t, U, k4, &q4);
        for (int i = 0; i < 4; ++i) {
            dual s = dadd(dadd(k1[i], dscale(2, k2[i])), dadd(dscale(2, k3[i]), k4[i]));
            X[i] = dadd(X[i], dscale(DT / 6, s));
        }
        dual sq = dadd(dadd(q1, dscale(2, q2)), dadd(dscale(2, q3), q4));
        *Q = dadd(*Q, dscale(DT / 6, sq));
    }
}

static void seed(const double x[4], const double u[6], dual X[4], dual U[6])
{
    for (int i = 0; i < 4; ++i) { X[i] = dconst(x[i]); X[i].d[i] = 1.0; }
    for (int i = 0; i < 6; ++i) { U[i] = dconst(u[i]); U[i].d[4 + i] = 1.0; }
}

void orc_f(const eepacc_params* p, double mass_scale, const double x[4], const double u[6],
           double xdot[4], double* L)
{
    dual X[4], U[6], xd[4], Lv;
    seed(x, u, X, U);
    f_dual(p, p->m * mass_scale, X, U, xd, &Lv);
    for (int i = 0; i < 4; ++i) xdot[i] = xd[i].v;
    *L = Lv.v;
}

void orc_F_jac(const eepacc_params* p, double mass_scale, const double x0[4], const double u[6],
               double xf[4], double* qf, double J[50])
{
    dual X[4], U[6], Q;
    seed(x0, u, X, U);
    F_dual(p, p->m * mass_scale, X, U, &Q);
    for (int i = 0; i < 4; ++i) xf[i] = X[i].v;
    *qf = Q.v;
    if (J) {
        for (int c = 0; c < 10; ++c) {
            for (int i = 0; i < 4; ++i) J[c * 5 + i] = X[i].d[c];
            J[c * 5 + 4] = Q.d[c];
        }
    }
}

void orc_F(const eepacc_params* p, double mass_scale, const double x0[4], const double u[6],
           double xf[4], double* qf)
{
    orc_F_jac(p, mass_scale, x0, u, xf, qf, 0);
}

/* ------------------------------------------------------------------------------------------
 * F/MPCs/EstimateVehicleTrajectory.m:55-88.  Tvec is uniform (= Ts) in every configuration
 * with golden data (Settings.m:120); estSetting: 0 ego AB/FB, 1 lead, 2 TVMPC, 3 BLMPC.
 * ------------------------------------------------------------------------------------------ */
void orc_est_traj(const eepacc_params* p, int estSetting, double s, double v, double a,
                  const double* sp, const double* vp, double* s_est, double* v_est)
{
    int mode;
    if (estSetting == 2) mode = p->est_tv;
    else if (estSetting == 3) mode = p->est_bl;
    else if (estSetting == 0) mode = p->est_ego;
    else mode = p->est_tar;
    const double tConst = (estSetting == 1) ? p->tConstACC_tar : p->tConstACC_ego;
    const double Ts = p->Ts;
    if (mode == 0) {
        s_est[0] = s;
        for (int i = 1; i <= N; ++i) s_est[i] = s_est[i - 1] + Ts * v;
        for (int i = 0; i <= N; ++i) v_est[i] = v;
    } else if (mode == 1) {
        s_est[0] = s; v_est[0] = v;
        for (int i = 1; i <= N; ++i) {          /* MATLAB i = 2..N+1, i.e. (i+1) here */
            if ((double)(i + 1) <= tConst / Ts && v_est[i - 1] + Ts * a > 0)
                v_est[i] = v_est[i - 1] + Ts * a;
            else
                v_est[i] = v_est[i - 1];
            s_est[i] = s_est[i - 1] + Ts * v_est[i - 1];
        }
    } else {
        s_est[0] = s; v_est[0] = v;
        for (int i = 1; i < N; ++i) { s_est[i] = sp[i + 1]; v_est[i] = vp[i + 1]; }
        s_est[N] = sp[N] + Ts * vp[N];
        v_est[N] = vp[N];
    }
}

/* ------------------------------------------------------------------------------------------
 * F/MPCs/EstimateRouteAndComfortBounds.m:60-199, including its quirks:
 *   :63-77   slope estimate is always slope(1)            :84  fallback v_lim = s_speedLim(end)
 *   :98      strict inequalities in the curve lookup      :121 phase uses t_0 + i*Tvec(i)
 * ------------------------------------------------------------------------------------------ */
static double mat_mod(double a, double m)
{
    if (m == 0.0) return a;
    double r = a - floor(a / m) * m;
    return r;
}

void orc_bounds(const eepacc_params* p, const eepacc_route* r, int mpctype,
                const double* s_est, const double* v_est, double t0,
                double* slope_est, double* v_lim, double* v_stop, double* v_tl, double* v_curv,
                double* a_min, double* a_max, double* j_min, double* j_max)
{
    for (int i = 0; i < N; ++i) {
        /* slope (:63-77): the inner loop always leaves at j == 1 */
        slope_est[i] = (r->n_slope > 0) ? r->slope[0] : 0.0;
        /* speed limit (:80-90) */
        v_lim[i] = 0.0;
        for (int j = 0; j < r->n_sl; ++j) {
            if (j == r->n_sl - 1) {
                v_lim[i] = r->s_speedLim[r->n_sl - 1];
            } else if (s_est[i] >= r->s_speedLim[j] && s_est[i] < r->s_speedLim[j + 1]) {
                v_lim[i] = r->v_speedLim[j];
                break;
            }
        }
        /* curve (:93-103) */
        v_curv[i] = 0.0;
        for (int j = 0; j < r->n_curv; ++j) {
            if (j == r->n_curv - 1) {
                v_curv[i] = p->alpha_TTL * pow(fabs(r->curvature[r->n_curv - 1]), -1.0 / 3.0);
            } else if (s_est[i] > r->s_curv[j] && s_est[i] < r->s_curv[j + 1]) {
                v_curv[i] = p->alpha_TTL * pow(fabs(r->curvature[j]), -1.0 / 3.0);
                break;
            }
        }
        /* stops (:106-114) */
        v_stop[i] = 1e5;
        for (int j = 0; j < r->n_stop; ++j) {
            double dist = fabs(r->stopLoc[j] - s_est[i]);
            if (dist < p->stopRefDist) v_stop[i] = dist * p->stopRefVelSlope + p->stopVel;
        }
        /* traffic lights (:117-134); MATLAB i = i+1 */
        v_tl[i] = 1e5;
        for (int j = 0; j < r->n_tl; ++j) {
            const double* tl = r->TLLoc[j];
            if (mat_mod(t0 + (double)(i + 1) * p->Ts - tl[1], tl[2] + tl[3]) < tl[2]) {
                double dist = tl[0] - s_est[i];
                if (fabs(dist) < p->stopRefDist) {
                    if (dist < 0) v_tl[i] = fabs(dist) * p->stopRefVelSlope + p->TLstopVel;
                    else if (fabs(dist) < p->TLStopRegionSize) v_tl[i] = p->TLstopVel;
                    else v_tl[i] = fabs(dist - p->stopVel) * p->stopRefVelSlope + p->TLstopVel;
                }
            }
        }
        /* comfort limits (:145-199) */
        double v = v_est[i];
        if (mpctype == 0) {
            if (v < 5) { a_min[i] = -5; a_max[i] = 4; j_min[i] = -5; j_max[i] = 5; }
            else if (v < 20) {
                a_min[i] = -5.5 + v / 10; a_max[i] = 14.0 / 3 - 2 * v / 15;
                j_min[i] = -35.0 / 6 + v / 6; j_max[i] = 35.0 / 6 - v / 6;
            } else { a_min[i] = -3.5; a_max[i] = 2; j_min[i] = -2.5; j_max[i] = 2.5; }
        } else {
            double aL, aH, jL, jH;
            if (mpctype == 1) { aL = p->BL_a_low; aH = p->BL_a_high; jL = p->BL_j_low; jH = p->BL_j_high; }
            else { aL = p->TV_a_low; aH = p->TV_a_high; jL = p->TV_j_low; jH = p->TV_j_high; }
            if (v < 5) { a_min[i] = -aL; a_max[i] = aL; j_min[i] = -jL; j_max[i] = jL; }
            else if (v < 20) {
                a_max[i] = (4 * aL - aH) / 3 + (aH - aL) / 15 * v; a_min[i] = -a_max[i];
                j_max[i] = (4 * jL - jH) / 3 + (jH - jL) / 15 * v; j_min[i] = -j_max[i];
            } else { a_min[i] = -aH; a_max[i] = aH; j_min[i] = -jH; j_max[i] = jH; }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * MC/RunOpt_ABMPC.m:288-323 (identical in RunOpt_BLMPC.m:261-296, RunOpt_TVMPC.m:237-272)
 * ------------------------------------------------------------------------------------------ */
void orc_force_alloc(const eepacc_params* p, const eepacc_route* r, double s_meas, double v_meas,
                     double a_cmd, double* Fm, double* Fb, double* a_real)
{
    const double m = p->m * r->mass_scale;
    double theta = orc_interp_pwa(s_meas, r->s_slope, r->slope, r->n_slope);
    double F_r = -p->zeta_a * v_meas * v_meas - p->c_r * m * p->g * cos(theta) - m * p->g * sin(theta);
    double F_t_req = m * p->lambda * a_cmd - F_r;
    double F_f_r_max = p->mu / p->L * (m * p->g * (p->L_f * cos(theta) + p->h_g * sin(theta)) +
                                       p->h_g * (p->zeta_a * v_meas * v_meas + p->lambda * m * a_cmd));
    double F_f_tot_max = p->mu * m * p->g * cos(theta);
    if (F_t_req < 0) {
        double F_m_min;
        if (v_meas < p->omega_m_r / p->phi) F_m_min = -p->phi * p->T_m_max / p->eta_TF;
        else F_m_min = -p->P_m_max / p->eta_TF / v_meas;
        double f = F_t_req;
        if (F_m_min > f) f = F_m_min;
        if (-F_f_r_max > f) f = -F_f_r_max;
        *Fm = f;
        *Fb = (F_t_req > -F_f_tot_max ? F_t_req : -F_f_tot_max) - f;
    } else {
        double F_m_max;
        if (v_meas < p->omega_m_r / p->phi) F_m_max = p->phi * p->T_m_max * p->eta_TF;
        else F_m_max = p->P_m_max * p->eta_TF / v_meas;
        double f = F_t_req;
        if (F_m_max < f) f = F_m_max;
        if (F_f_r_max < f) f = F_f_r_max;
        *Fm = f;
        *Fb = 0.0;
    }
    *a_real = (*Fm + *Fb + F_r) / m / p->lambda;
}

/* ------------------------------------------------------------------------------------------
 * Sparse-ordered QP in z = [x0;u0;x1;u1;...;xN] (CreateQP_*.m, solverToUse==1 branches).
 * Hs is kept dense (nz x nz, symmetric); constraint rows as short (index,value) lists.
 * ------------------------------------------------------------------------------------------ */
#define ROWNZ 6
typedef struct {
    int nz, nxu, nu, nc;
    double Hs[ORC_NZMAX * ORC_NZMAX];
    double cs[ORC_NZMAX];
    int    ri[ORC_MMAX][ROWNZ];
    double rv[ORC_MMAX][ROWNZ];
    int    rn[ORC_MMAX];
    double lb[ORC_MMAX], ub[ORC_MMAX];
} sparse_qp;

static void sq_init(sparse_qp* q, int nu)
{
    q->nu = nu; q->nxu = 2 + nu; q->nz = q->nxu * N + 2; q->nc = 0;
    memset(q->Hs, 0, sizeof(double) * (size_t)q->nz * q->nz);
    memset(q->cs, 0, sizeof(double) * (size_t)q->nz);
}
static void addH(sparse_qp* q, const int* ii, int n, const double* M, double scale)
{
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b)
            q->Hs[(size_t)ii[b] * q->nz + ii[a]] += scale * M[a * n + b];
}
static void row(sparse_qp* q, int cnt, const int* idx, const double* val, double lb, double ub)
{
    int c = q->nc++;
    q->rn[c] = cnt;
    for (int k = 0; k < cnt; ++k) { q->ri[c][k] = idx[k]; q->rv[c][k] = val[k]; }
    q->lb[c] = lb; q->ub[c] = ub;
}
#define ROW1(a, va, lb, ub) do { int i_[1] = { a }; double v_[1] = { va }; row(q, 1, i_, v_, lb, ub); } while (0)
#define ROW2(a, va, b, vb, lb, ub) do { int i_[2] = { a, b }; double v_[2] = { va, vb }; row(q, 2, i_, v_, lb, ub); } while (0)
#define ROW3(a, va, b, vb, c, vc, lb, ub) do { int i_[3] = { a, b, c }; double v_[3] = { va, vb, vc }; row(q, 3, i_, v_, lb, ub); } while (0)
#define ROW5(a, va, b, vb, c, vc, d, vd, e, ve, lb, ub) do { int i_[5] = { a, b, c, d, e }; double v_[5] = { va, vb, vc, vd, ve }; row(q, 5, i_, v_, lb, ub); } while (0)

typedef struct {
    double slope[N], v_lim[N], v_stop[N], v_tl[N], v_curv[N], a_min[N], a_max[N], j_min[N], j_max[N];
} bounds_t;

/* F/MPCs/CreateQP_FB.m:56-489 (qpOASES dense branch, FBuseTaylor = true, Mb == 0) */
static void create_qp_fb(const eepacc_params* p, const eepacc_route* r, double s_0, double v_0,
                         const double* s_est, const double* v_est, const double* s_tv_est,
                         double t0, double a_minus1, sparse_qp* q, double* theta_est)
{
    const double m = p->m * r->mass_scale;
    const double* W = p->W_FB;
    const double w_P = W[0], w_a = W[1], w_j = W[2], w_v = W[3], w_h = W[4], w_s = W[5], w_f = W[6];
    const double* b = p->b_quadr;
    const double INF = INFINITY;
    bounds_t B;
    orc_bounds(p, r, 0, s_est, v_est, t0, B.slope, B.v_lim, B.v_stop, B.v_tl, B.v_curv,
               B.a_min, B.a_max, B.j_min, B.j_max);
    for (int i = 0; i < N; ++i) theta_est[i] = B.slope[i];
    sq_init(q, 6);
    const double s_min = s_0, s_max = p->s_goal, v_min = 0, v_max = p->v_max;
    const double Fm_min = -1e4, Fm_max = 1e4, Fb_min = -1e4, Fb_max = 0;
    const double za = p->zeta_a, lm = p->lambda * m, cphi = (30.0 / PI_) * p->phi;
    double zeta_rg = 0, Dzeta_rg = 0;
    for (int kk = 0; kk < N; ++kk) {
        const int k = kk;                /* 0-based estimate index (MATLAB k = kk+1) */
        const double Tp = p->Ts;
        const int o = kk * 8;
        const int sc = o, vc = o + 1, Fmc = o + 2, Fbc = o + 3, xv = o + 4, xh = o + 5, xs = o + 6, xf = o + 7;
        const int vp = vc - 8, Fmp = Fmc - 8, Fbp = Fbc - 8;
        if (kk > 0) {
            double prev = zeta_rg;
            zeta_rg = m * p->g * (p->c_r * cos(theta_est[k]) + sin(theta_est[k]));
            Dzeta_rg = zeta_rg - prev;
        } else {
            zeta_rg = m * p->g * (p->c_r * cos(theta_est[k]) + sin(theta_est[k]));
        }
        /* power (:181-184) */
        {
            int ii[2] = { vc, Fmc };
            double M[4] = { 2 * cphi * cphi * b[5], cphi * b[4], cphi * b[4], 2 * b[3] };
            addH(q, ii, 2, M, w_P);
            q->cs[vc] += w_P * cphi * b[2];
            q->cs[Fmc] += w_P * b[1];
        }
        /* acceleration penalty (:187-191) */
        {
            int ii[3] = { vc, Fmc, Fbc };
            double ve = v_est[k];
            double M[9] = { za * za * ve * ve + za * zeta_rg, -za * ve, -za * ve,
                            -za * ve, 1, 1,
                            -za * ve, 1, 1 };
            addH(q, ii, 3, M, 2 * w_a / (lm * lm));
            q->cs[Fmc] += w_a / (lm * lm) * (-2 * zeta_rg);
            q->cs[Fbc] += w_a / (lm * lm) * (-2 * zeta_rg);
        }
        /* jerk penalty (:194-208) */
        if (kk == 0) {
            int ii[2] = { Fmc, Fbc };
            double M[4] = { 1, 1, 1, 1 };
            double sc2 = 2 * w_j / ((lm * Tp) * (lm * Tp));
            addH(q, ii, 2, M, sc2);
            double cc = 2 * w_j * (za * v_0 * v_0 + m * p->g * (p->c_r * cos(theta_est[0]) + sin(theta_est[0])) + lm * a_minus1)
                        / ((lm * Tp) * (lm * Tp));
            q->cs[Fmc] -= cc;
            q->cs[Fbc] -= cc;
        } else {
            int ii[6] = { vp, Fmp, Fbp, vc, Fmc, Fbc };
            double vk = v_est[k], vk1 = v_est[k - 1];
            double M[36] = {
                za * za * vk * vk + 2 * za * Dzeta_rg, -za * vk1, -za * vk1, -za * za * vk * vk1, za * vk1, za * vk1,
                -za * vk1, 1, 1, za * vk, -1, -1,
                -za * vk1, 1, 1, za * vk, -1, -1,
                -za * za * vk * vk1, za * vk, za * vk, za * za * vk1 * vk1 - 2 * za * Dzeta_rg, -za * vk, -za * vk,
                za * vk1, -1, -1, -za * vk, 1, 1,
                za * vk1, -1, -1, -za * vk, 1, 1 };
            double sc2 = 2 * w_j / ((lm * Tp) * (lm * Tp));
            addH(q, ii, 6, M, sc2);
            q->cs[Fmp] += sc2 * Dzeta_rg;
            q->cs[Fbp] += sc2 * Dzeta_rg;
            q->cs[Fmc] -= sc2 * Dzeta_rg;
            q->cs[Fbc] -= sc2 * Dzeta_rg;
        }
        /* slacks (:211-215) */
        q->cs[xv] += w_v;
        q->cs[xh] += 1e2 * w_h;
        q->Hs[(size_t)xh * q->nz + xh] += 2 * w_h;
        q->cs[xs] += w_s;
        q->cs[xf] += w_f;
        /* simple bounds as rows (:311-342) */
        ROW1(sc, 1, s_min, s_max);
        ROW1(vc, 1, v_min, v_max);
        ROW1(Fmc, 1, Fm_min, Fm_max);
        ROW1(Fbc, 1, Fb_min, Fb_max);
        ROW1(xv, 1, 0, INF);
        ROW1(xh, 1, 0, INF);
        ROW1(xs, 1, 0, INF);
        ROW1(xf, 1, 0, INF);
        /* torque limits (:359-366) */
        ROW3(vc, -p->phi * p->T_m_max * p->T_m_max / 4 / p->P_m_max, Fmc, p->eta_TF / p->phi, xf, 1, -p->T_m_max, INF);
        ROW3(vc, p->phi * p->T_m_max * p->T_m_max / 4 / p->P_m_max, Fmc, 1 / p->eta_TF / p->phi, xf, -1, -INF, p->T_m_max);
        /* rear wheel traction (:369-377) */
        double zeta_w = m * p->g * (p->L_f * cos(theta_est[k]) + p->h_g * sin(theta_est[k]));
        ROW3(Fmc, p->L / p->mu + p->h_g, Fbc, p->h_g, xf, 1, -zeta_w + p->h_g * zeta_rg, INF);
        ROW3(Fmc, p->L / p->mu - p->h_g, Fbc, -p->h_g, xf, -1, -INF, zeta_w - p->h_g * zeta_rg);
        /* total friction (:380-387) */
        ROW3(Fmc, 1, Fbc, 1, xf, -1, -INF, p->mu * m * p->g * cos(theta_est[k]));
        ROW3(Fmc, 1, Fbc, 1, xf, 1, -p->mu * m * p->g * cos(theta_est[k]), INF);
        /* ISO acceleration (:390-397) */
        double vv = v_est[k] * v_est[k];
        ROW3(Fmc, 1, Fbc, 1, xf, -1, -INF, lm * B.a_max[k] + za * vv + zeta_rg);
        ROW3(Fmc, 1, Fbc, 1, xf, 1, lm * B.a_min[k] + za * vv + zeta_rg, INF);
        /* ISO jerk (:400-418) */
        if (kk == 0) {
            ROW3(Fmc, 1, Fbc, 1, xf, -1, -INF, lm * (Tp * B.j_max[k] + a_minus1) + za * vv + zeta_rg);
            ROW3(Fmc, 1, Fbc, 1, xf, 1, lm * (Tp * B.j_min[k] + a_minus1) + za * vv + zeta_rg, INF);
        } else {
            double dv2 = za * (vv - v_est[k - 1] * v_est[k - 1]);
            ROW5(Fmp, -1, Fbp, -1, Fmc, 1, Fbc, 1, xf, -1, -INF, lm * Tp * B.j_max[k] + dv2 + Dzeta_rg);
            ROW5(Fmp, -1, Fbp, -1, Fmc, 1, Fbc, 1, xf, 1, lm * Tp * B.j_min[k] + dv2 + Dzeta_rg, INF);
        }
        /* route (:421-442) */
        ROW2(vc, 1, xf, -1, -INF, B.v_lim[k]);
        ROW2(vc, 1, xf, -1, -INF, B.v_curv[k]);
        ROW2(vc, 1, xs, -1, -INF, B.v_stop[k]);
        ROW2(vc, 1, xs, -1, -INF, B.v_tl[k]);
        /* incentive (:445-448) */
        ROW2(vc, 1, xv, 1, (B.v_lim[k] < B.v_curv[k] ? B.v_lim[k] : B.v_curv[k]), INF);
        /* safe headway (:451-458) */
        ROW2(sc, 1, xs, -1, -INF, s_tv_est[k] - p->h_min);
        ROW3(sc, 1, vc, p->tau_min, xs, -1, -INF, s_tv_est[k]);
        /* desired headway, Taylor form (:461-467) */
        ROW3(sc, 1, vc, p->T_hwp + 2 * p->G_hwp * v_est[k], xh, -1, -INF,
             s_tv_est[k] - p->A_hwp + p->G_hwp * vv);
    }
    /* final stage (:478-489); note s_tv_est(N) (1-based) = index N-1 */
    {
        const int o = N * 8;
        ROW1(o, 1, -INF, s_tv_est[N - 1] - p->h_min);
        ROW2(o, 1, o + 1, p->tau_min, -INF, s_tv_est[N - 1]);
    }
}

/* F/MPCs/CreateQP_AB.m:49-371 (MC) + ABO/Functions/MPCs/CreateQP_AB.m:152-166,324-347 */
static void create_qp_ab(const eepacc_params* p, const eepacc_route* r, double s_0, double v_0,
                         const double* s_est, const double* v_est, const double* s_tv_est,
                         double t0, double a_minus1, sparse_qp* q)
{
    (void)s_0; (void)v_0;
    const double m = p->m * r->mass_scale;
    const double* W = p->W_AB;
    const double w_FC = W[0], w_a = W[1], w_j = W[2], w_v = W[3], w_h = W[4], w_s = W[5], w_f = W[6];
    const double INF = INFINITY;
    bounds_t B;
    orc_bounds(p, r, 0, s_est, v_est, t0, B.slope, B.v_lim, B.v_stop, B.v_tl, B.v_curv,
               B.a_min, B.a_max, B.j_min, B.j_max);
    sq_init(q, 5);
    const double s_min = 0, s_max = p->s_goal, v_min = 0, v_max = p->v_max;
    for (int kk = 0; kk < N; ++kk) {
        const int k = kk;
        const double T = p->Ts;
        const int o = kk * 7;
        const int sc = o, vc = o + 1, ac = o + 2, xv = o + 3, xh = o + 4, xs = o + 5, xf = o + 6;
        const int ap = ac - 7;
        /* fuel term, ABO tree only (ABO CreateQP_AB.m:161-166); w_FC = 0 for the MC tree */
        if (w_FC != 0.0) {
            q->Hs[(size_t)vc * q->nz + vc] += w_FC * 2 * p->p01 * p->F2;
            q->cs[vc] += w_FC * p->p10;
            q->cs[ac] += w_FC * p->p01 * p->lambda * m;
        }
        /* acceleration (:151-153) */
        q->Hs[(size_t)ac * q->nz + ac] += 2 * w_a;
        /* jerk (:156-163) */
        if (kk == 0) {
            q->Hs[(size_t)ac * q->nz + ac] += 2 * w_j / (T * T);
            q->cs[ac] -= 2 * w_j / T * a_minus1;
        } else {
            int ii[2] = { ac, ap };
            double M[4] = { 1, -1, -1, 1 };
            addH(q, ii, 2, M, 2 * w_j / (T * T));
        }
        /* slacks (:166-170) */
        q->cs[xv] += w_v;
        q->cs[xh] += 1e2 * w_h;
        q->Hs[(size_t)xh * q->nz + xh] += 2 * w_h;
        q->cs[xs] += w_s;
        q->cs[xf] += w_f;
        /* bounds (:239-262) */
        ROW1(sc, 1, s_min, s_max);
        ROW1(vc, 1, v_min, v_max);
        ROW1(xv, 1, 0, INF);
        ROW1(xh, 1, 0, INF);
        ROW1(xs, 1, 0, INF);
        ROW1(xf, 1, 0, INF);
        /* ISO acceleration (:275-282) */
        ROW2(ac, 1, xf, -1, -INF, B.a_max[k]);
        ROW2(ac, 1, xf, 1, B.a_min[k], INF);
        /* ISO jerk (:285-305) */
        if (kk > 0) {
            ROW3(ap, -1, ac, 1, xf, -1, -INF, T * B.j_max[k]);
            ROW3(ap, -1, ac, 1, xf, 1, T * B.j_min[k], INF);
        } else {
            ROW2(ac, 1, xf, -1, -INF, T * B.j_max[k] + a_minus1);
            ROW2(ac, 1, xf, 1, T * B.j_min[k] + a_minus1, INF);
        }
        /* route rows: MC tree only (MC CreateQP_AB.m:307-329; commented out in ABO) */
        if (p->ab_route_rows) {
            ROW2(vc, 1, xf, -1, -INF, B.v_lim[k]);
            ROW2(vc, 1, xf, -1, -INF, B.v_curv[k]);
            ROW2(vc, 1, xs, -1, -INF, B.v_stop[k]);
            ROW2(vc, 1, xs, -1, -INF, B.v_tl[k]);
        }
        /* incentive (:332-335) */
        ROW2(vc, 1, xv, 1, (B.v_lim[k] < B.v_curv[k] ? B.v_lim[k] : B.v_curv[k]), INF);
        /* safe headway (:338-345) */
        ROW2(sc, 1, xs, -1, -INF, s_tv_est[k] - p->h_min);
        ROW3(sc, 1, vc, p->tau_min, xs, -1, -INF, s_tv_est[k]);
        /* desired headway (:348-354) */
        ROW3(sc, 1, vc, p->T_hwp + p->G_hwp * v_est[k], xh, -1, -INF, s_tv_est[k] - p->A_hwp);
    }
    {
        const int o = N * 7;
        ROW1(o, 1, -INF, s_tv_est[N - 1] - p->h_min);
        ROW2(o, 1, o + 1, p->tau_min, -INF, s_tv_est[N - 1]);
    }
}

/* F/MPCs/CreateQP_BL.m:46-314 and CreateQP_TV.m (diff: :44-45 v_est=0, :265,271 0.8x, no headway) */
static void create_qp_bltv(int tv, const eepacc_params* p, const eepacc_route* r,
                           const double* s_est, const double* v_est_in, const double* s_tv_est,
                           double t0, double a_minus1, sparse_qp* q)
{
    const double* W = tv ? p->W_TV : p->W_BL;
    const double w_v = W[0], w_a = W[1], w_j = W[2], w_f = W[3];
    const double INF = INFINITY;
    const double Ts = p->Ts;
    double v_zero[N + 1];
    memset(v_zero, 0, sizeof v_zero);
    const double* v_est = tv ? v_zero : v_est_in;
    bounds_t B;
    orc_bounds(p, r, tv ? 2 : 1, s_est, v_est, t0, B.slope, B.v_lim, B.v_stop, B.v_tl, B.v_curv,
               B.a_min, B.a_max, B.j_min, B.j_max);
    sq_init(q, 2);
    const double sf = tv ? 0.8 : 1.0;
    for (int kk = 0; kk < N; ++kk) {
        const int k = kk;
        const int o = kk * 4;
        const int sc = o, vc = o + 1, ac = o + 2, xf = o + 3, ap = ac - 4;
        q->cs[vc] -= w_v;
        q->Hs[(size_t)ac * q->nz + ac] += 2 * w_a;
        if (kk == 0) {
            q->Hs[(size_t)ac * q->nz + ac] += 2 * w_j / (Ts * Ts);
            q->cs[ac] -= 2 * w_j / Ts * a_minus1;
        } else {
            int ii[2] = { ac, ap };
            double M[4] = { 1, -1, -1, 1 };
            addH(q, ii, 2, M, 2 * w_j / (Ts * Ts));
        }
        q->cs[xf] += w_f;
        ROW1(sc, 1, 0, p->s_goal);
        ROW1(vc, 1, 0, p->v_max);
        ROW1(xf, 1, 0, INF);
        ROW2(ac, 1, xf, 1, B.a_min[k], INF);
        ROW2(ac, 1, xf, -1, -INF, B.a_max[k]);
        if (kk > 0) {
            ROW3(ap, -1, ac, 1, xf, 1, Ts * B.j_min[k], INF);
            ROW3(ap, -1, ac, 1, xf, -1, -INF, Ts * B.j_max[k]);
        } else {
            ROW2(ac, 1, xf, 1, Ts * B.j_min[k] + a_minus1, INF);
            ROW2(ac, 1, xf, -1, -INF, Ts * B.j_max[k] + a_minus1);
        }
        ROW2(vc, 1, xf, -1, -INF, sf * B.v_lim[k]);
        ROW2(vc, 1, xf, -1, -INF, sf * B.v_curv[k]);
        ROW2(vc, 1, xf, -1, -INF, B.v_stop[k]);
        ROW2(vc, 1, xf, -1, -INF, B.v_tl[k]);
        if (!tv) {
            ROW2(sc, 1, xf, -1, -INF, s_tv_est[k] - p->h_min);
            ROW3(sc, 1, vc, p->tau_min, xf, -1, -INF, s_tv_est[k]);
        }
    }
    {
        const int o = N * 4;
        q->cs[o + 1] -= w_v;
        if (!tv) {
            ROW1(o, 1, -INF, s_tv_est[N - 1] - p->h_min);
            ROW2(o, 1, o + 1, p->tau_min, -INF, s_tv_est[N - 1]);
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * F/MPCs/TransformToDenseFormulation.m:30-91
 * A[k] 2x2 row-major, Bc[k] = the one non-zero 2-vector column of B (same for every driven
 * control column), nb = number of driven control columns (AB/BL/TV: 1 (a); FB: 2 (Fm, Fb)).
 * ------------------------------------------------------------------------------------------ */
static void condense(const sparse_qp* q, const double A[][4], const double Bc[][2], int nb,
                     const double D[][2], double s0, double v0, orc_qp* out)
{
    const int nu = q->nu, nxu = q->nxu, nz = q->nz, n = N * nu, m = q->nc;
    out->n = n; out->m = m; out->nz = nz;
    double* Psi = out->Psi;      /* col-major nz x n */
    double* d = out->d;
    memset(Psi, 0, sizeof(double) * (size_t)nz * n);
    memset(d, 0, sizeof(double) * (size_t)nz);
    /* control rows: identity (:78-82) */
    for (int k = 0; k < N; ++k)
        for (int j = 0; j < nu; ++j) Psi[(size_t)(k * nu + j) * nz + (k * nxu + 2 + j)] = 1.0;
    /* state rows (:43-68) */
    for (int i = 0; i < N; ++i) {
        double M0 = Bc[i][0], M1 = Bc[i][1];
        for (int k = i + 1; k <= N; ++k) {
            for (int j = 0; j < nb; ++j) {
                Psi[(size_t)(i * nu + j) * nz + (k * nxu)] = M0;
                Psi[(size_t)(i * nu + j) * nz + (k * nxu + 1)] = M1;
            }
            if (k < N) {
                double n0 = A[k][0] * M0 + A[k][1] * M1, n1 = A[k][2] * M0 + A[k][3] * M1;
                M0 = n0; M1 = n1;
            }
        }
    }
    {
        double x0 = s0, x1 = v0;
        d[0] = x0; d[1] = x1;
        for (int k = 1; k <= N; ++k) {
            double n0 = A[k - 1][0] * x0 + A[k - 1][1] * x1 + D[k - 1][0];
            double n1 = A[k - 1][2] * x0 + A[k - 1][3] * x1 + D[k - 1][1];
            x0 = n0; x1 = n1;
            d[k * nxu] = x0; d[k * nxu + 1] = x1;
        }
    }
    /* H = Psi' Hs Psi ; c = Psi'(0.5 (Hs+Hs') d + cs)   (:87-88).  Hs is banded. */
    static double HP[ORC_NZMAX * ORC_NMAX];
    const int bw = 2 * nxu;
    for (int j = 0; j < n; ++j) {
        for (int a = 0; a < nz; ++a) {
            double s = 0;
            int lo = a - bw < 0 ? 0 : a - bw, hi = a + bw >= nz ? nz - 1 : a + bw;
            for (int b = lo; b <= hi; ++b) s += q->Hs[(size_t)b * nz + a] * Psi[(size_t)j * nz + b];
            HP[(size_t)j * nz + a] = s;
        }
    }
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
            double s = 0;
            for (int a = 0; a < nz; ++a) {
                double pv = Psi[(size_t)i * nz + a];
                if (pv != 0.0) s += pv * HP[(size_t)j * nz + a];
            }
            out->H[(size_t)j * n + i] = s;
        }
    {
        double t[ORC_NZMAX];
        for (int a = 0; a < nz; ++a) {
            double s = 0;
            for (int b = 0; b < nz; ++b)
                s += 0.5 * (q->Hs[(size_t)b * nz + a] + q->Hs[(size_t)a * nz + b]) * d[b];
            t[a] = s + q->cs[a];
        }
        for (int i = 0; i < n; ++i) {
            double s = 0;
            for (int a = 0; a < nz; ++a) s += Psi[(size_t)i * nz + a] * t[a];
            out->g[i] = s;
        }
    }
    /* G = Gs Psi ; bounds shifted by Gs d (:89-91) */
    for (int c = 0; c < m; ++c) {
        double gd = 0;
        for (int e = 0; e < q->rn[c]; ++e) gd += q->rv[c][e] * d[q->ri[c][e]];
        out->lba[c] = q->lb[c] - gd;
        out->uba[c] = q->ub[c] - gd;
    }
    for (int j = 0; j < n; ++j)
        for (int c = 0; c < m; ++c) {
            double s = 0;
            for (int e = 0; e < q->rn[c]; ++e) s += q->rv[c][e] * Psi[(size_t)j * nz + q->ri[c][e]];
            out->A[(size_t)j * m + c] = s;
        }
}

/* MC/RunOpt_FBMPC.m:78-90 with v_0 = v_init */
void orc_fb_ad_init(const eepacc_params* p, double mass_scale, orc_fb_ad* ad)
{
    const double m = p->m * mass_scale, T = p->Ts, v0 = p->v_init;
    for (int k = 0; k < N; ++k) {
        ad->A[k][0] = 1; ad->A[k][1] = T; ad->A[k][2] = 0;
        ad->A[k][3] = 1 - 2 * T * p->zeta_a * v0 * v0 / p->lambda / m;
        ad->D[k][0] = 0;
        ad->D[k][1] = T / p->lambda / m * (p->zeta_a * v0 * v0);
    }
}

int orc_qp_nvar(int ctrl) { return ctrl == EEPACC_FB ? 120 : ctrl == EEPACC_AB ? 100 : 40; }
int orc_qp_ncon(int ctrl, const eepacc_params* p)
{
    switch (ctrl) {
    case EEPACC_FB: return 522;
    case EEPACC_AB: return p->ab_route_rows ? 362 : 282;
    case EEPACC_BL: return 262;
    default: return 220;
    }
}

int orc_qpbuild(int ctrl, const eepacc_params* p, const eepacc_route* r,
                double s0, double v0, double a_minus1, double s_tv, double v_tv, double a_tv,
                double t0, const double* s_prev_sol, const double* v_prev_sol,
                orc_fb_ad* ad, int k_sim, orc_qp* qp)
{
    static sparse_qp q;   /* oracle is single-threaded per process */
    double s_est[N + 1], v_est[N + 1], s_tv_est[N + 1], v_tv_est[N + 1], theta_est[N];
    double A[N][4], Bc[N][2], D[N][2];
    const double m = p->m * r->mass_scale, T = p->Ts;
    int est = ctrl == EEPACC_BL ? 3 : ctrl == EEPACC_TV ? 2 : 0;
    orc_est_traj(p, est, s0, v0, a_minus1, s_prev_sol, v_prev_sol, s_est, v_est);
    if (ctrl != EEPACC_TV) orc_est_traj(p, 1, s_tv, v_tv, a_tv, 0, 0, s_tv_est, v_tv_est);
    if (ctrl == EEPACC_FB) {
        create_qp_fb(p, r, s0, v0, s_est, v_est, s_tv_est, t0, a_minus1, &q, theta_est);
        /* RunOpt_FBMPC.m:247-259 */
        if (p->fb_ad_quirk) {
            int row = k_sim - 1;             /* A(k,..): k is the SIMULATION step */
            if (row < N) {
                int i = N - 1;               /* the loop's last pass wins */
                ad->A[row][0] = 1; ad->A[row][1] = T; ad->A[row][2] = 0;
                ad->A[row][3] = 1 - 2 * T * p->zeta_a * v_est[i] / p->lambda / m;
                ad->D[row][0] = 0;
                ad->D[row][1] = T / p->lambda / m * (p->zeta_a * v_est[i] * v_est[i] -
                                m * p->g * (p->c_r * cos(theta_est[i]) + sin(theta_est[i])));
            }
        } else {
            for (int i = 0; i < N; ++i) {
                ad->A[i][0] = 1; ad->A[i][1] = T; ad->A[i][2] = 0;
                ad->A[i][3] = 1 - 2 * T * p->zeta_a * v_est[i] / p->lambda / m;
                ad->D[i][0] = 0;
                ad->D[i][1] = T / p->lambda / m * (p->zeta_a * v_est[i] * v_est[i] -
                              m * p->g * (p->c_r * cos(theta_est[i]) + sin(theta_est[i])));
            }
        }
        for (int k = 0; k < N; ++k) {
            memcpy(A[k], ad->A[k], sizeof A[k]);
            D[k][0] = ad->D[k][0]; D[k][1] = ad->D[k][1];
            Bc[k][0] = 0; Bc[k][1] = T / p->lambda / m;      /* RunOpt_FBMPC.m:89 */
        }
        condense(&q, A, Bc, 2, D, s0, v0, qp);
    } else {
        if (ctrl == EEPACC_AB) create_qp_ab(p, r, s0, v0, s_est, v_est, s_tv_est, t0, a_minus1, &q);
        else create_qp_bltv(ctrl == EEPACC_TV, p, r, s_est, v_est, s_tv_est, t0, a_minus1, &q);
        for (int k = 0; k < N; ++k) {                         /* RunOpt_ABMPC.m:76-81 */
            A[k][0] = 1; A[k][1] = T; A[k][2] = 0; A[k][3] = 1;
            Bc[k][0] = .5 * T * T; Bc[k][1] = T;
            D[k][0] = 0; D[k][1] = 0;
        }
        condense(&q, A, Bc, 1, D, s0, v0, qp);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Closed loop: one iteration of `for kk = 0:N_sim` of RunOpt_{AB,FB,BL,TV}MPC.m
 * ------------------------------------------------------------------------------------------ */
struct orc_vehicle {
    int ctrl, kk;
    eepacc_params p;
    eepacc_route r;
    double s_prev, v_prev, Fm_prev, Fb_prev, a_minus1, v_tv_meas, t0;
    double s_prev_sol[N + 1], v_prev_sol[N + 1];
    orc_fb_ad ad;
    orc_qp qp;
    orc_qpws* ws;
    double x[ORC_NMAX], lam_a[ORC_MMAX], lam_x[ORC_NMAX];
};

orc_vehicle* orc_vehicle_new(int ctrl, const eepacc_params* p, const eepacc_route* r)
{
    orc_vehicle* v = (orc_vehicle*)calloc(1, sizeof *v);
    if (!v) return 0;
    v->ctrl = ctrl; v->p = *p; v->r = *r;
    v->a_minus1 = p->a_minus1_init;
    orc_fb_ad_init(p, r->mass_scale, &v->ad);
    v->ws = orc_qpws_new(orc_qp_nvar(ctrl), orc_qp_ncon(ctrl, p));
    return v;
}
void orc_vehicle_free(orc_vehicle* v) { if (v) { orc_qpws_free(v->ws); free(v); } }
const orc_qp* orc_vehicle_qp(const orc_vehicle* v) { return &v->qp; }
void orc_vehicle_get_ws(const orc_vehicle* v, int* ws) { orc_qpws_get_ws(v->ws, ws); }
/* last closed-loop QP solution: x [n], lam_a [m] (conic scheme signs) */
void orc_vehicle_get_sol(const orc_vehicle* v, double* x, double* lam_a)
{
    for (int i = 0; i < v->qp.n; ++i) x[i] = v->x[i];
    for (int i = 0; i < v->qp.m; ++i) lam_a[i] = v->lam_a[i];
}

int orc_vehicle_step(orc_vehicle* V, double s_tv_k, double v_tv_k, double* out)
{
    const eepacc_params* p = &V->p;
    const eepacc_route* r = &V->r;
    const int ctrl = V->ctrl;
    const double Ts = p->Ts;
    double s_meas, v_meas, a_tv_prev;
    if (V->kk == 0) {                               /* RunOpt_FBMPC.m:166-183 */
        s_meas = p->s_init; v_meas = p->v_init;
        V->v_tv_meas = 0.0;
        a_tv_prev = 0.0;
    } else {                                        /* :185-201 */
        double x0[2] = { V->s_prev, V->v_prev }, u[2] = { V->Fm_prev, V->Fb_prev }, xn[2];
        orc_plant(p, r, x0, u, xn);
        s_meas = xn[0]; v_meas = xn[1];
        V->a_minus1 = (v_meas - V->v_prev) / Ts;
        double v_tv_prev = V->v_tv_meas;
        V->v_tv_meas = v_tv_k;
        a_tv_prev = (V->v_tv_meas - v_tv_prev) / Ts;
    }
    orc_qpbuild(ctrl, p, r, s_meas, v_meas, V->a_minus1, s_tv_k, V->v_tv_meas, a_tv_prev, V->t0,
                V->s_prev_sol, V->v_prev_sol, &V->ad, V->kk + 1, &V->qp);
    double cost = 0; int iter = 0;
    int status = orc_qpsolve(V->ws, V->qp.H, V->qp.g, V->qp.A, V->qp.lba, V->qp.uba, 0, 0,
                             V->x, &cost, V->lam_a, V->lam_x, &iter);
    /* z = Psi x + d (RunOpt_FBMPC.m:287) */
    const orc_qp* q = &V->qp;
    const int nu = q->n / N, nxu = nu + 2;
    double z[ORC_NZMAX];
    for (int a = 0; a < q->nz; ++a) {
        double s = q->d[a];
        for (int j = 0; j < q->n; ++j) {
            double pv = q->Psi[(size_t)j * q->nz + a];
            if (pv != 0.0) s += pv * V->x[j];
        }
        z[a] = s;
    }
    for (int k = 0; k <= N; ++k) { V->s_prev_sol[k] = z[k * nxu]; V->v_prev_sol[k] = z[k * nxu + 1]; }
    double s_opt = z[0], v_opt = z[1], Fm, Fb, a_out, xi[4] = { 0, 0, 0, 0 };
    if (ctrl == EEPACC_FB) {
        Fm = z[2]; Fb = z[3]; xi[0] = z[4]; xi[1] = z[5]; xi[2] = z[6]; xi[3] = z[7];
        a_out = (V->kk > 0) ? (v_opt - V->v_prev) / Ts : 0.0;   /* :316-318 */
    } else {
        double a_cmd = z[2];
        if (ctrl == EEPACC_AB) { xi[0] = z[3]; xi[1] = z[4]; xi[2] = z[5]; xi[3] = z[6]; }
        else xi[3] = z[3];
        orc_force_alloc(p, r, s_meas, v_meas, a_cmd, &Fm, &Fb, &a_out);
    }
    V->s_prev = s_opt; V->v_prev = v_opt; V->Fm_prev = Fm; V->Fb_prev = Fb;
    V->t0 += Ts;
    V->kk += 1;
    out[0] = s_opt; out[1] = v_opt; out[2] = Fm; out[3] = Fb;
    out[4] = xi[0]; out[5] = xi[1]; out[6] = xi[2]; out[7] = xi[3];
    out[8] = a_out; out[9] = (double)status; out[10] = (double)iter; out[11] = cost;
    return 0;
}

/* KPI block (Main.m:136-148,218-258; RunOpt_FBMPC.m:334-339):
 * [0] E_opt(end) J  [1] s_opt(end) m  [2] travel time at cut-off s  [3] E at cut-off J
 * [4] max a  [5] min a  [6] rms a (1:distInd)  [7] failed QPs */
void orc_kpi(const eepacc_params* p, double mass_scale, const double* log, int n, double* kpi)
{
    (void)mass_scale;
    int distInd = 0;                      /* 1-based like Main.m */
    for (int i = 1; i <= n - 1; ++i) {
        if (log[(i - 1) * 12] < p->cutOffDist && log[i * 12] > p->cutOffDist) { distInd = i; break; }
    }
    if (distInd == 0) distInd = n - 1;
    double E = 0, Ecut = 0, amax = -INFINITY, amin = INFINITY, a2 = 0, fail = 0;
    for (int i = 1; i <= n; ++i) {
        const double* o = log + (size_t)(i - 1) * 12;
        double rpm = (30.0 / PI_) * o[1] * p->phi;
        E += p->Ts * orc_power5(p->b_fifth, o[2], rpm);
        if (i == distInd - 1) Ecut = E;
        if (i <= distInd) {
            if (o[8] > amax) amax = o[8];
            if (o[8] < amin) amin = o[8];
            a2 += o[8] * o[8];
        }
        if (o[9] != 0.0) fail += 1;
    }
    kpi[0] = E; kpi[1] = log[(size_t)(n - 1) * 12];
    kpi[2] = 0.1 * floor(distInd * p->Ts * 10 + 0.5);
    kpi[3] = Ecut; kpi[4] = amax; kpi[5] = amin; kpi[6] = sqrt(a2 / distInd); kpi[7] = fail;
    /* comfort: j = diff(a)/Ts (RunOpt_FBMPC.m:339), max / min / rms over j(1:distInd) (Main.m:247-250) */
    {
        double jmax = 0, jmin = 0, j2 = 0;
        int nj = distInd < n - 1 ? distInd : n - 1;
        for (int i = 1; i <= nj; ++i) {
            double j = (log[(size_t)i * 12 + 8] - log[(size_t)(i - 1) * 12 + 8]) / p->Ts;
            if (i == 1 || j > jmax) jmax = j;
            if (i == 1 || j < jmin) jmin = j;
            j2 += j * j;
        }
        kpi[8] = jmax; kpi[9] = jmin; kpi[10] = nj > 0 ? sqrt(j2 / nj) : 0.0;
    }
    /* constraint violations: steps with an active soft constraint (xi_v, xi_h, xi_s, xi_f; CreateQP_FB.m:211-215) */
    for (int q = 0; q < 4; ++q) {
        double cnt = 0;
        for (int i = 0; i < n; ++i) if (log[(size_t)i * 12 + 4 + q] > 1e-6) cnt += 1;
        kpi[11 + q] = cnt;
    }
    kpi[15] = 0; kpi[16] = 0; kpi[17] = INFINITY;      /* headway figures need the lead trace: orc_run */
}

int orc_run(int ctrl, const eepacc_params* p, const eepacc_route* r, const double* lead,
            int lead_stride, int n_steps, double* log, double* kpi)
{
    orc_vehicle* v = orc_vehicle_new(ctrl, p, r);
    if (!v) return 1;
    for (int k = 0; k < n_steps; ++k)
        orc_vehicle_step(v, lead[(size_t)k * lead_stride], lead[(size_t)k * lead_stride + 1],
                         log + (size_t)k * 12);
    if (kpi) {
        orc_kpi(p, r->mass_scale, log, n_steps, kpi);
        if (ctrl != EEPACC_TV) {               /* headway against the lead trace (the TVMPC has no target vehicle) */
            for (int k = 0; k < n_steps; ++k) {
                double gap = lead[(size_t)k * lead_stride] - log[(size_t)k * 12];
                if (!isfinite(gap)) continue;
                if (gap < p->h_min) kpi[15] += 1;
                if (gap < p->tau_min * log[(size_t)k * 12 + 1]) kpi[16] += 1;
                if (gap < kpi[17]) kpi[17] = gap;
            }
        }
    }
    orc_vehicle_free(v);
    return 0;
}

/* B scenarios one after the other (the C-side loop of the CPU baseline: no per-scenario interpreter overhead).
 * routes[B]; lead [n_steps][B][2] (the engine's layout); kpi [B][EEPACC_N_KPI]; log_scratch [n_steps][12]. */
int orc_run_batch(int ctrl, const eepacc_params* p, const eepacc_route* routes, const double* lead, long long B,
                  int n_steps, double* kpi, double* log_scratch)
{
    for (long long b = 0; b < B; ++b) {
        int rc = orc_run(ctrl, p, routes + b, lead + 2 * b, (int)(2 * B), n_steps, log_scratch,
                         kpi ? kpi + b * EEPACC_N_KPI : 0);
        if (rc) return rc;
    }
    return 0;
}
