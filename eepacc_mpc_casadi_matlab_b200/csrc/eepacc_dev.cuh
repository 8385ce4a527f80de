// eepacc_dev.cuh -- device-side building blocks shared by the sm_100a kernels.
//
// Every function cites the reference code it replaces:
//   MC/  = MATLAB_CasADi/        ABO/ = ACCMPC-ABO_CasADi/        F/ = <tree>/Functions/
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "../../include/eepacc.h"

#define EEP_N EEPACC_N_HOR
#define EEP_PI 3.14159265358979323846

// F/PWA_function_manipulation/InterpPWA.m:14-27
__device__ __forceinline__ double eep_interp_pwa(double d, const double* __restrict__ doms,
                                                 const double* __restrict__ vals, int n)
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
    return vals[n - 1];
}

// F/Other/GetMotorPower_FifthOrderSurface.m:16-20
__device__ __forceinline__ double eep_power5(const double* __restrict__ b, double x, double y)
{
    double x2 = x * x, x3 = x2 * x, x4 = x3 * x, x5 = x4 * x;
    double y2 = y * y, y3 = y2 * y, y4 = y3 * y, y5 = y4 * y;
    return b[0] + b[1] * x + b[2] * y + b[3] * x2 + b[4] * x * y + b[5] * y2 +
           b[6] * x3 + b[7] * x2 * y + b[8] * x * y2 + b[9] * y3 + b[10] * x4 +
           b[11] * x3 * y + b[12] * x2 * y2 + b[13] * x * y3 + b[14] * y4 +
           b[15] * x5 + b[16] * x4 * y + b[17] * x3 * y2 + b[18] * x2 * y3 +
           b[19] * x * y4 + b[20] * y5;
}

// F/Other/GetMotorPower_QuadraticFit.m:13
__device__ __forceinline__ double eep_power2(const double* __restrict__ b, double Fm, double rpm)
{
    return b[0] + b[1] * Fm + b[2] * rpm + b[3] * Fm * Fm + b[4] * Fm * rpm + b[5] * rpm * rpm;
}

// F/MPCs/RunPlantModel.m:41-44 (f_NL)
__device__ __forceinline__ void eep_plant_rhs(const eepacc_params& p, const eepacc_route* __restrict__ r,
                                              double m, double s, double v, double u, double& ds, double& dv)
{
    double theta = eep_interp_pwa(s, r->s_slope, r->slope, r->n_slope);
    double sn, cs;
    sincos(theta, &sn, &cs);
    ds = v;
    dv = 1.0 / p.lambda / m * (u - p.zeta_a * v * v - p.c_r * m * p.g * cs - m * p.g * sn);
}

// F/MPCs/RunPlantModel.m:20-37 : M RK4 sub-steps
__device__ __forceinline__ void eep_plant_step(const eepacc_params& p, const eepacc_route* __restrict__ r,
                                               double& s, double& v, double Fm, double Fb)
{
    const int M = p.n_integrate_plant;
    const double m = p.m * r->mass_scale;
    const double u = Fm + Fb;
    const double DT = p.Ts / M;
    for (int k = 0; k < M; ++k) {
        double k1s, k1v, k2s, k2v, k3s, k3v, k4s, k4v;
        eep_plant_rhs(p, r, m, s, v, u, k1s, k1v);
        eep_plant_rhs(p, r, m, s + DT / 2 * k1s, v + DT / 2 * k1v, u, k2s, k2v);
        eep_plant_rhs(p, r, m, s + DT / 2 * k2s, v + DT / 2 * k2v, u, k3s, k3v);
        eep_plant_rhs(p, r, m, s + DT * k3s, v + DT * k3v, u, k4s, k4v);
        s = s + DT / 6 * (k1s + 2 * k2s + 2 * k3s + k4s);
        v = v + DT / 6 * (k1v + 2 * k2v + 2 * k3v + k4v);
    }
}

// MC/RunOpt_ABMPC.m:288-323 (same block in RunOpt_BLMPC.m:261-296, RunOpt_TVMPC.m:237-272)
__device__ __forceinline__ void eep_force_alloc(const eepacc_params& p, const eepacc_route* __restrict__ r,
                                                double s_meas, double v_meas, double a_cmd,
                                                double& Fm, double& Fb, double& a_real)
{
    const double m = p.m * r->mass_scale;
    double theta = eep_interp_pwa(s_meas, r->s_slope, r->slope, r->n_slope);
    double sn, cs;
    sincos(theta, &sn, &cs);
    double F_r = -p.zeta_a * v_meas * v_meas - p.c_r * m * p.g * cs - m * p.g * sn;
    double F_t_req = m * p.lambda * a_cmd - F_r;
    double F_f_r_max = p.mu / p.L * (m * p.g * (p.L_f * cs + p.h_g * sn) +
                                     p.h_g * (p.zeta_a * v_meas * v_meas + p.lambda * m * a_cmd));
    double F_f_tot_max = p.mu * m * p.g * cs;
    if (F_t_req < 0) {
        double F_m_min = (v_meas < p.omega_m_r / p.phi) ? -p.phi * p.T_m_max / p.eta_TF
                                                        : -p.P_m_max / p.eta_TF / v_meas;
        double f = fmax(fmax(F_t_req, F_m_min), -F_f_r_max);
        Fm = f;
        Fb = fmax(F_t_req, -F_f_tot_max) - f;
    } else {
        double F_m_max = (v_meas < p.omega_m_r / p.phi) ? p.phi * p.T_m_max * p.eta_TF
                                                        : p.P_m_max * p.eta_TF / v_meas;
        Fm = fmin(fmin(F_t_req, F_m_max), F_f_r_max);
        Fb = 0.0;
    }
    a_real = (Fm + Fb + F_r) / m / p.lambda;
}

// F/MPCs/EstimateVehicleTrajectory.m:55-88 (uniform Tvec = Ts, Settings.m:120)
__device__ __forceinline__ void eep_est_traj(const eepacc_params& p, int estSetting, double s, double v,
                                             double a, const double* sp, const double* vp,
                                             double* s_est, double* v_est)
{
    int mode = estSetting == 2 ? p.est_tv : estSetting == 3 ? p.est_bl : estSetting == 0 ? p.est_ego : p.est_tar;
    const double tConst = (estSetting == 1) ? p.tConstACC_tar : p.tConstACC_ego;
    const double Ts = p.Ts;
    if (mode == 0) {
        s_est[0] = s;
        for (int i = 1; i <= EEP_N; ++i) s_est[i] = s_est[i - 1] + Ts * v;
        for (int i = 0; i <= EEP_N; ++i) v_est[i] = v;
    } else if (mode == 1) {
        s_est[0] = s; v_est[0] = v;
        for (int i = 1; i <= EEP_N; ++i) {
            if ((double)(i + 1) <= tConst / Ts && v_est[i - 1] + Ts * a > 0) v_est[i] = v_est[i - 1] + Ts * a;
            else v_est[i] = v_est[i - 1];
            s_est[i] = s_est[i - 1] + Ts * v_est[i - 1];
        }
    } else {
        s_est[0] = s; v_est[0] = v;
        for (int i = 1; i < EEP_N; ++i) { s_est[i] = sp[i + 1]; v_est[i] = vp[i + 1]; }
        s_est[EEP_N] = sp[EEP_N] + Ts * vp[EEP_N];
        v_est[EEP_N] = vp[EEP_N];
    }
}

__device__ __forceinline__ double eep_mat_mod(double a, double m)
{
    if (m == 0.0) return a;
    return a - floor(a / m) * m;
}

// Per-stage route and comfort bounds: F/MPCs/EstimateRouteAndComfortBounds.m:60-199 (one stage i,
// 0-based).  Quirks kept: slope always slope(1) (:63-77), fallback v_lim = s_speedLim(end) (:84),
// strict curve lookup (:98), light phase at t_0 + i*Tvec(i) (:121).
struct eep_stage_bounds {
    double slope, v_lim, v_stop, v_tl, v_curv, a_min, a_max, j_min, j_max;
};

__device__ __forceinline__ void eep_bounds_stage(const eepacc_params& p, const eepacc_route* __restrict__ r,
                                                 int mpctype, int i, double s_est, double v_est, double t0,
                                                 eep_stage_bounds& B)
{
    B.slope = (r->n_slope > 0) ? r->slope[0] : 0.0;
    B.v_lim = 0.0;
    for (int j = 0; j < r->n_sl; ++j) {
        if (j == r->n_sl - 1) {
            B.v_lim = r->s_speedLim[r->n_sl - 1];
        } else if (s_est >= r->s_speedLim[j] && s_est < r->s_speedLim[j + 1]) {
            B.v_lim = r->v_speedLim[j];
            break;
        }
    }
    B.v_curv = 0.0;
    for (int j = 0; j < r->n_curv; ++j) {
        if (j == r->n_curv - 1) {
            B.v_curv = p.alpha_TTL * pow(fabs(r->curvature[r->n_curv - 1]), -1.0 / 3.0);
        } else if (s_est > r->s_curv[j] && s_est < r->s_curv[j + 1]) {
            B.v_curv = p.alpha_TTL * pow(fabs(r->curvature[j]), -1.0 / 3.0);
            break;
        }
    }
    B.v_stop = 1e5;
    for (int j = 0; j < r->n_stop; ++j) {
        double dist = fabs(r->stopLoc[j] - s_est);
        if (dist < p.stopRefDist) B.v_stop = dist * p.stopRefVelSlope + p.stopVel;
    }
    B.v_tl = 1e5;
    for (int j = 0; j < r->n_tl; ++j) {
        const double* tl = r->TLLoc[j];
        if (eep_mat_mod(t0 + (double)(i + 1) * p.Ts - tl[1], tl[2] + tl[3]) < tl[2]) {
            double dist = tl[0] - s_est;
            if (fabs(dist) < p.stopRefDist) {
                if (dist < 0) B.v_tl = fabs(dist) * p.stopRefVelSlope + p.TLstopVel;
                else if (fabs(dist) < p.TLStopRegionSize) B.v_tl = p.TLstopVel;
                else B.v_tl = fabs(dist - p.stopVel) * p.stopRefVelSlope + p.TLstopVel;
            }
        }
    }
    double v = v_est;
    if (mpctype == 0) {
        if (v < 5) { B.a_min = -5; B.a_max = 4; B.j_min = -5; B.j_max = 5; }
        else if (v < 20) {
            B.a_min = -5.5 + v / 10; B.a_max = 14.0 / 3 - 2 * v / 15;
            B.j_min = -35.0 / 6 + v / 6; B.j_max = 35.0 / 6 - v / 6;
        } else { B.a_min = -3.5; B.a_max = 2; B.j_min = -2.5; B.j_max = 2.5; }
    } else {
        double aL, aH, jL, jH;
        if (mpctype == 1) { aL = p.BL_a_low; aH = p.BL_a_high; jL = p.BL_j_low; jH = p.BL_j_high; }
        else { aL = p.TV_a_low; aH = p.TV_a_high; jL = p.TV_j_low; jH = p.TV_j_high; }
        if (v < 5) { B.a_min = -aL; B.a_max = aL; B.j_min = -jL; B.j_max = jL; }
        else if (v < 20) {
            B.a_max = (4 * aL - aH) / 3 + (aH - aL) / 15 * v; B.a_min = -B.a_max;
            B.j_max = (4 * jL - jH) / 3 + (jH - jL) / 15 * v; B.j_min = -B.j_max;
        } else { B.a_min = -aH; B.a_max = aH; B.j_min = -jH; B.j_max = jH; }
    }
}
