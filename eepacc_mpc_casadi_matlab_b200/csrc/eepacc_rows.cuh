// eepacc_rows.cuh -- the constraint rows of the four MPCs in structured form.
//
// One row of the sparse-ordered QP touches at most (s_k, v_k) of ONE stage k plus <= 3 controls of
// stage k and, for the jerk rows, the driven controls of stage k-1 (CreateQP_FB.m:310-489,
// CreateQP_AB.m:237-371, CreateQP_BL.m:215-314, CreateQP_TV.m).  K2 uses the full description
// (coefficients + bounds) to write the dense conic operands; K3's matrix-free path only needs the
// coefficients (BOUNDS = false), which depend on nothing but the parameters and v_est(k).
//
// SRC must provide v_est[k]; with BOUNDS also bnd[k], zrg[k], s_tv_est[k], s_meas, a_minus1.
#pragma once
#include "eepacc_dev.cuh"

// Plain scalar slots (no indexed arrays): the descriptor must live in registers -- K3 runs with the
// maximum shared-memory carve-out, so anything that lands in local memory pays an L2 round trip.
struct RowDesc {
    double cS, cV;      // coefficients on s_k, v_k
    int kst;            // stage of the state entries (-1: none, -2: none + "-1 on the driven controls of stage k-1")
    int nctl;           // 1..3 control entries
    int cj0, cj1, cj2;  // dense column index of control entries
    double cv0, cv1, cv2;
    double lb, ub;
};

__device__ __forceinline__ void rd_clear(RowDesc& r)
{
    r.cS = 0; r.cV = 0; r.kst = -1; r.nctl = 0; r.lb = -INFINITY; r.ub = INFINITY;
    r.cj0 = r.cj1 = r.cj2 = -1; r.cv0 = r.cv1 = r.cv2 = 0.0;
}
__device__ __forceinline__ void rd_ctl(RowDesc& r, int j, double v)          // first entry
{ r.cj0 = j; r.cv0 = v; r.nctl = 1; }
__device__ __forceinline__ void rd_ctl2(RowDesc& r, int j, double v)         // second entry
{ r.cj1 = j; r.cv1 = v; r.nctl = 2; }
__device__ __forceinline__ void rd_ctl3(RowDesc& r, int j, double v)         // third entry
{ r.cj2 = j; r.cv2 = v; r.nctl = 3; }

// CreateQP_FB.m:310-473 (26 rows / stage) and :478-489 (2 terminal rows)
template <bool BOUNDS, class SRC>
__device__ void fb_row(const eepacc_params& p, double m, const SRC& S, int c, RowDesc& r)
{
    rd_clear(r);
    const int NH = EEP_N;
    if (c >= 26 * NH) {
        r.kst = NH; r.cS = 1;
        if (c == 26 * NH) { if constexpr (BOUNDS) r.ub = S.s_tv_est[NH - 1] - p.h_min; }
        else { r.cV = p.tau_min; if constexpr (BOUNDS) r.ub = S.s_tv_est[NH - 1]; }
        return;
    }
    const int k = c / 26, t = c % 26;
    const int jFm = k * 6, jFb = jFm + 1, jxv = jFm + 2, jxh = jFm + 3, jxs = jFm + 4, jxf = jFm + 5;
    const double ve = S.v_est[k];
    switch (t) {
    case 0: r.kst = k; r.cS = 1; if constexpr (BOUNDS) { r.lb = S.s_meas; r.ub = p.s_goal; } break;            // s  (:311)
    case 1: r.kst = k; r.cV = 1; if constexpr (BOUNDS) { r.lb = 0; r.ub = p.v_max; } break;                    // v
    case 2: rd_ctl(r, jFm, 1); if constexpr (BOUNDS) { r.lb = -1e4; r.ub = 1e4; } break;                       // Fm
    case 3: rd_ctl(r, jFb, 1); if constexpr (BOUNDS) { r.lb = -1e4; r.ub = 0; } break;                         // Fb
    case 4: rd_ctl(r, jxv, 1); if constexpr (BOUNDS) r.lb = 0; break;
    case 5: rd_ctl(r, jxh, 1); if constexpr (BOUNDS) r.lb = 0; break;
    case 6: rd_ctl(r, jxs, 1); if constexpr (BOUNDS) r.lb = 0; break;
    case 7: rd_ctl(r, jxf, 1); if constexpr (BOUNDS) r.lb = 0; break;
    case 8:  // torque limit, min (:359)
        r.kst = k; r.cV = -p.phi * p.T_m_max * p.T_m_max / 4 / p.P_m_max;
        rd_ctl(r, jFm, p.eta_TF / p.phi); rd_ctl2(r, jxf, 1);
        if constexpr (BOUNDS) r.lb = -p.T_m_max;
        break;
    case 9:  // torque limit, max (:363)
        r.kst = k; r.cV = p.phi * p.T_m_max * p.T_m_max / 4 / p.P_m_max;
        rd_ctl(r, jFm, 1 / p.eta_TF / p.phi); rd_ctl2(r, jxf, -1);
        if constexpr (BOUNDS) r.ub = p.T_m_max;
        break;
    case 10: // rear wheel traction (:369-372)
        rd_ctl(r, jFm, p.L / p.mu + p.h_g); rd_ctl2(r, jFb, p.h_g); rd_ctl3(r, jxf, 1);
        if constexpr (BOUNDS) {
            double sn, cn; sincos(S.bnd[k].slope, &sn, &cn);
            double zw = m * p.g * (p.L_f * cn + p.h_g * sn);
            r.lb = -zw + p.h_g * S.zrg[k];
        }
        break;
    case 11: // (:374-376)
        rd_ctl(r, jFm, p.L / p.mu - p.h_g); rd_ctl2(r, jFb, -p.h_g); rd_ctl3(r, jxf, -1);
        if constexpr (BOUNDS) {
            double sn, cn; sincos(S.bnd[k].slope, &sn, &cn);
            double zw = m * p.g * (p.L_f * cn + p.h_g * sn);
            r.ub = zw - p.h_g * S.zrg[k];
        }
        break;
    case 12: rd_ctl(r, jFm, 1); rd_ctl2(r, jFb, 1); rd_ctl3(r, jxf, -1);                                         // (:380)
        if constexpr (BOUNDS) { double sn, cn; sincos(S.bnd[k].slope, &sn, &cn); r.ub = p.mu * m * p.g * cn; }
        break;
    case 13: rd_ctl(r, jFm, 1); rd_ctl2(r, jFb, 1); rd_ctl3(r, jxf, 1);                                          // (:384)
        if constexpr (BOUNDS) { double sn, cn; sincos(S.bnd[k].slope, &sn, &cn); r.lb = -p.mu * m * p.g * cn; }
        break;
    case 14: rd_ctl(r, jFm, 1); rd_ctl2(r, jFb, 1); rd_ctl3(r, jxf, -1);                                         // ISO a (:390)
        if constexpr (BOUNDS) r.ub = p.lambda * m * S.bnd[k].a_max + p.zeta_a * ve * ve + S.zrg[k];
        break;
    case 15: rd_ctl(r, jFm, 1); rd_ctl2(r, jFb, 1); rd_ctl3(r, jxf, 1);
        if constexpr (BOUNDS) r.lb = p.lambda * m * S.bnd[k].a_min + p.zeta_a * ve * ve + S.zrg[k];
        break;
    case 16: case 17: {   // ISO jerk (:400-418)
        const double sg = (t == 16) ? -1.0 : 1.0;
        rd_ctl(r, jFm, 1); rd_ctl2(r, jFb, 1); rd_ctl3(r, jxf, sg);
        if (k > 0) r.kst = -2;                  // -1 on (Fm, Fb) of stage k-1
        if constexpr (BOUNDS) {
            const double lm = p.lambda * m, za = p.zeta_a, Tp = p.Ts, vv = ve * ve;
            const double jb = (t == 16) ? S.bnd[k].j_max : S.bnd[k].j_min;
            double val;
            if (k == 0) val = lm * (Tp * jb + S.a_minus1) + za * vv + S.zrg[k];
            else val = lm * Tp * jb + za * (vv - S.v_est[k - 1] * S.v_est[k - 1]) + (S.zrg[k] - S.zrg[k - 1]);
            if (t == 16) r.ub = val; else r.lb = val;
        }
        break; }
    case 18: r.kst = k; r.cV = 1; rd_ctl(r, jxf, -1); if constexpr (BOUNDS) r.ub = S.bnd[k].v_lim; break;      // (:421)
    case 19: r.kst = k; r.cV = 1; rd_ctl(r, jxf, -1); if constexpr (BOUNDS) r.ub = S.bnd[k].v_curv; break;
    case 20: r.kst = k; r.cV = 1; rd_ctl(r, jxs, -1); if constexpr (BOUNDS) r.ub = S.bnd[k].v_stop; break;
    case 21: r.kst = k; r.cV = 1; rd_ctl(r, jxs, -1); if constexpr (BOUNDS) r.ub = S.bnd[k].v_tl; break;
    case 22: r.kst = k; r.cV = 1; rd_ctl(r, jxv, 1);                                                           // (:445)
        if constexpr (BOUNDS) r.lb = fmin(S.bnd[k].v_lim, S.bnd[k].v_curv);
        break;
    case 23: r.kst = k; r.cS = 1; rd_ctl(r, jxs, -1); if constexpr (BOUNDS) r.ub = S.s_tv_est[k] - p.h_min; break;   // (:451)
    case 24: r.kst = k; r.cS = 1; r.cV = p.tau_min; rd_ctl(r, jxs, -1); if constexpr (BOUNDS) r.ub = S.s_tv_est[k]; break;
    default: // 25: desired headway, Taylor form (:464-467)
        r.kst = k; r.cS = 1; r.cV = p.T_hwp + 2 * p.G_hwp * ve; rd_ctl(r, jxh, -1);
        if constexpr (BOUNDS) r.ub = S.s_tv_est[k] - p.A_hwp + p.G_hwp * ve * ve;
        break;
    }
}

// CreateQP_AB.m:237-371 (MC: 18 rows / stage, ABO: 14)
template <bool BOUNDS, class SRC>
__device__ void ab_row(const eepacc_params& p, const SRC& S, int c, RowDesc& r)
{
    rd_clear(r);
    const int NH = EEP_N;
    const int rps = p.ab_route_rows ? 18 : 14;
    if (c >= rps * NH) {
        r.kst = NH; r.cS = 1;
        if (c == rps * NH) { if constexpr (BOUNDS) r.ub = S.s_tv_est[NH - 1] - p.h_min; }
        else { r.cV = p.tau_min; if constexpr (BOUNDS) r.ub = S.s_tv_est[NH - 1]; }
        return;
    }
    const int k = c / rps;
    int t = c % rps;
    if (!p.ab_route_rows && t >= 10) t += 4;       // skip the four route rows
    const int ja = k * 5, jxv = ja + 1, jxh = ja + 2, jxs = ja + 3, jxf = ja + 4;
    const double T = p.Ts;
    switch (t) {
    case 0: r.kst = k; r.cS = 1; if constexpr (BOUNDS) { r.lb = 0; r.ub = p.s_goal; } break;
    case 1: r.kst = k; r.cV = 1; if constexpr (BOUNDS) { r.lb = 0; r.ub = p.v_max; } break;
    case 2: rd_ctl(r, jxv, 1); if constexpr (BOUNDS) r.lb = 0; break;
    case 3: rd_ctl(r, jxh, 1); if constexpr (BOUNDS) r.lb = 0; break;
    case 4: rd_ctl(r, jxs, 1); if constexpr (BOUNDS) r.lb = 0; break;
    case 5: rd_ctl(r, jxf, 1); if constexpr (BOUNDS) r.lb = 0; break;
    case 6: rd_ctl(r, ja, 1); rd_ctl2(r, jxf, -1); if constexpr (BOUNDS) r.ub = S.bnd[k].a_max; break;          // (:275)
    case 7: rd_ctl(r, ja, 1); rd_ctl2(r, jxf, 1); if constexpr (BOUNDS) r.lb = S.bnd[k].a_min; break;
    case 8: case 9: {                                                                                          // (:285-305)
        const double sg = (t == 8) ? -1.0 : 1.0;
        rd_ctl(r, ja, 1); rd_ctl2(r, jxf, sg);
        if (k > 0) r.kst = -2;
        if constexpr (BOUNDS) {
            const double jb = (t == 8) ? S.bnd[k].j_max : S.bnd[k].j_min;
            const double val = (k > 0) ? T * jb : T * jb + S.a_minus1;
            if (t == 8) r.ub = val; else r.lb = val;
        }
        break; }
    case 10: r.kst = k; r.cV = 1; rd_ctl(r, jxf, -1); if constexpr (BOUNDS) r.ub = S.bnd[k].v_lim; break;      // MC :307-329
    case 11: r.kst = k; r.cV = 1; rd_ctl(r, jxf, -1); if constexpr (BOUNDS) r.ub = S.bnd[k].v_curv; break;
    case 12: r.kst = k; r.cV = 1; rd_ctl(r, jxs, -1); if constexpr (BOUNDS) r.ub = S.bnd[k].v_stop; break;
    case 13: r.kst = k; r.cV = 1; rd_ctl(r, jxs, -1); if constexpr (BOUNDS) r.ub = S.bnd[k].v_tl; break;
    case 14: r.kst = k; r.cV = 1; rd_ctl(r, jxv, 1);                                                           // (:332)
        if constexpr (BOUNDS) r.lb = fmin(S.bnd[k].v_lim, S.bnd[k].v_curv);
        break;
    case 15: r.kst = k; r.cS = 1; rd_ctl(r, jxs, -1); if constexpr (BOUNDS) r.ub = S.s_tv_est[k] - p.h_min; break;
    case 16: r.kst = k; r.cS = 1; r.cV = p.tau_min; rd_ctl(r, jxs, -1); if constexpr (BOUNDS) r.ub = S.s_tv_est[k]; break;
    default: // 17 (:348-354)
        r.kst = k; r.cS = 1; r.cV = p.T_hwp + p.G_hwp * S.v_est[k]; rd_ctl(r, jxh, -1);
        if constexpr (BOUNDS) r.ub = S.s_tv_est[k] - p.A_hwp;
        break;
    }
}

// CreateQP_BL.m:215-314 (13 rows / stage + 2) ; CreateQP_TV.m (11 rows / stage, 0.8x limits)
template <bool BOUNDS, class SRC>
__device__ void bl_row(const eepacc_params& p, int tv, const SRC& S, int c, RowDesc& r)
{
    rd_clear(r);
    const int NH = EEP_N;
    const int rps = tv ? 11 : 13;
    if (c >= rps * NH) {
        r.kst = NH; r.cS = 1;
        if (c == rps * NH) { if constexpr (BOUNDS) r.ub = S.s_tv_est[NH - 1] - p.h_min; }
        else { r.cV = p.tau_min; if constexpr (BOUNDS) r.ub = S.s_tv_est[NH - 1]; }
        return;
    }
    const int k = c / rps, t = c % rps;
    const int ja = k * 2, jxf = ja + 1;
    const double Ts = p.Ts, sf = tv ? 0.8 : 1.0;
    switch (t) {
    case 0: r.kst = k; r.cS = 1; if constexpr (BOUNDS) { r.lb = 0; r.ub = p.s_goal; } break;
    case 1: r.kst = k; r.cV = 1; if constexpr (BOUNDS) { r.lb = 0; r.ub = p.v_max; } break;
    case 2: rd_ctl(r, jxf, 1); if constexpr (BOUNDS) r.lb = 0; break;
    case 3: rd_ctl(r, ja, 1); rd_ctl2(r, jxf, 1); if constexpr (BOUNDS) r.lb = S.bnd[k].a_min; break;           // (:232)
    case 4: rd_ctl(r, ja, 1); rd_ctl2(r, jxf, -1); if constexpr (BOUNDS) r.ub = S.bnd[k].a_max; break;
    case 5: case 6: {                                                                                          // (:242-262)
        const double sg = (t == 5) ? 1.0 : -1.0;
        rd_ctl(r, ja, 1); rd_ctl2(r, jxf, sg);
        if (k > 0) r.kst = -2;
        if constexpr (BOUNDS) {
            const double jb = (t == 5) ? S.bnd[k].j_min : S.bnd[k].j_max;
            const double val = (k > 0) ? Ts * jb : Ts * jb + S.a_minus1;
            if (t == 5) r.lb = val; else r.ub = val;
        }
        break; }
    case 7: r.kst = k; r.cV = 1; rd_ctl(r, jxf, -1); if constexpr (BOUNDS) r.ub = sf * S.bnd[k].v_lim; break;  // (:265)
    case 8: r.kst = k; r.cV = 1; rd_ctl(r, jxf, -1); if constexpr (BOUNDS) r.ub = sf * S.bnd[k].v_curv; break;
    case 9: r.kst = k; r.cV = 1; rd_ctl(r, jxf, -1); if constexpr (BOUNDS) r.ub = S.bnd[k].v_stop; break;
    case 10: r.kst = k; r.cV = 1; rd_ctl(r, jxf, -1); if constexpr (BOUNDS) r.ub = S.bnd[k].v_tl; break;
    case 11: r.kst = k; r.cS = 1; rd_ctl(r, jxf, -1); if constexpr (BOUNDS) r.ub = S.s_tv_est[k] - p.h_min; break;   // (:289)
    default: r.kst = k; r.cS = 1; r.cV = p.tau_min; rd_ctl(r, jxf, -1); if constexpr (BOUNDS) r.ub = S.s_tv_est[k]; break;
    }
}

template <bool BOUNDS, class SRC>
__device__ __forceinline__ void eep_row(int ctrl, const eepacc_params& p, double mass, const SRC& S, int c, RowDesc& r)
{
    if (ctrl == EEPACC_FB) fb_row<BOUNDS>(p, mass, S, c, r);
    else if (ctrl == EEPACC_AB) ab_row<BOUNDS>(p, S, c, r);
    else bl_row<BOUNDS>(p, ctrl == EEPACC_TV, S, c, r);
}
