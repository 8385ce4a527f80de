// k2_build.cuh -- device code of K5+K4+K2 for ONE vehicle (see k2_build.cu for the description); shared by the
// per-step kernel (k2_build.cu) and the fused closed-loop kernel (k3r_qp.cu: k5_cycle_kernel).
#pragma once
#include "eepacc_dev.cuh"
#include "eepacc_kernels.h"
#include "eepacc_rows.cuh"

#define N EEP_N
#define MAXW 7          // 1 + n_u (FB)


struct BuildSmem {
    double s_est[N + 1], v_est[N + 1], s_tv_est[N + 1], v_tv_est[N + 1];
    eep_stage_bounds bnd[N];
    double zrg[N];
    double alpha[N], Dv[N];                 // A_k(2,2), D_k(2)
    double PS[N + 1][N], PV[N + 1][N];      // state sensitivities to the driven controls of stage i
    double ds[N + 1], dv[N + 1];            // free response
    double Dk[N][MAXW * MAXW];
    double Ck[N][MAXW * MAXW];              // rows: stage k, cols: stage k-1
    double cs[N][MAXW];
    double cs_vN;                           // linear cost on v_N (BL/TV terminal)
    double tv[N + 1];                       // t(v_k) = (Hs d + cs)(v_k)
    double MV[N][EEPACC_MAX_NVAR];          // (Hs Psi)(v_k, j)
    double s_meas, v_meas, a_minus1, s_tv, v_tv, a_tv, t0;
    alignas(16) eepacc_route route;         // this vehicle's route preview, staged once per step (K5)
};

__device__ __forceinline__ void add3(double* D, int W, const int* ii, const double* M, int n, double sc)
{
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b) D[ii[a] * W + ii[b]] += sc * M[a * n + b];
}

// ---- FB: stage cost (CreateQP_FB.m:178-215); w = [v Fm Fb xi_v xi_h xi_s xi_f] ----
static __device__ void fb_jerk_block(const eepacc_params& p, double m, const BuildSmem& S, int k, double* M /*6x6*/,
                              double& sc2, double& Dz)
{
    const double za = p.zeta_a, lm = p.lambda * m, Tp = p.Ts;
    Dz = S.zrg[k] - S.zrg[k - 1];
    double vk = S.v_est[k], vk1 = S.v_est[k - 1];
    double Mloc[36] = {
        za * za * vk * vk + 2 * za * Dz, -za * vk1, -za * vk1, -za * za * vk * vk1, za * vk1, za * vk1,
        -za * vk1, 1, 1, za * vk, -1, -1,
        -za * vk1, 1, 1, za * vk, -1, -1,
        -za * za * vk * vk1, za * vk, za * vk, za * za * vk1 * vk1 - 2 * za * Dz, -za * vk, -za * vk,
        za * vk1, -1, -1, -za * vk, 1, 1,
        za * vk1, -1, -1, -za * vk, 1, 1 };
    for (int i = 0; i < 36; ++i) M[i] = Mloc[i];
    sc2 = 2 * p.W_FB[2] / ((lm * Tp) * (lm * Tp));
}

static __device__ void fb_stage_cost(const eepacc_params& p, double m, BuildSmem& S, int k)
{
    const int W = 7;
    const double* Wt = p.W_FB;
    const double w_P = Wt[0], w_a = Wt[1], w_j = Wt[2];
    const double* b = p.b_quadr;
    const double za = p.zeta_a, lm = p.lambda * m, cphi = (30.0 / EEP_PI) * p.phi, Tp = p.Ts;
    double* D = S.Dk[k]; double* C = S.Ck[k]; double* c = S.cs[k];
    for (int i = 0; i < W * W; ++i) { D[i] = 0; C[i] = 0; }
    for (int i = 0; i < W; ++i) c[i] = 0;
    const double zrg = S.zrg[k], ve = S.v_est[k];
    {   // power (:181-184)
        int ii[2] = { 0, 1 };
        double M[4] = { 2 * cphi * cphi * b[5], cphi * b[4], cphi * b[4], 2 * b[3] };
        add3(D, W, ii, M, 2, w_P);
        c[0] += w_P * cphi * b[2];
        c[1] += w_P * b[1];
    }
    {   // acceleration (:187-191)
        int ii[3] = { 0, 1, 2 };
        double M[9] = { za * za * ve * ve + za * zrg, -za * ve, -za * ve, -za * ve, 1, 1, -za * ve, 1, 1 };
        add3(D, W, ii, M, 3, 2 * w_a / (lm * lm));
        c[1] += w_a / (lm * lm) * (-2 * zrg);
        c[2] += w_a / (lm * lm) * (-2 * zrg);
    }
    if (k == 0) {   // jerk, first stage (:194-198)
        int ii[2] = { 1, 2 };
        double M[4] = { 1, 1, 1, 1 };
        double sc2 = 2 * w_j / ((lm * Tp) * (lm * Tp));
        add3(D, W, ii, M, 2, sc2);
        double sn, cn; sincos(S.bnd[0].slope, &sn, &cn);
        double cc = 2 * w_j * (za * S.v_meas * S.v_meas + m * p.g * (p.c_r * cn + sn) + lm * S.a_minus1)
                    / ((lm * Tp) * (lm * Tp));
        c[1] -= cc; c[2] -= cc;
    } else {        // jerk block of stage k: current-current part and coupling (:200-207)
        double M[36], sc2, Dz;
        fb_jerk_block(p, m, S, k, M, sc2, Dz);
        for (int a = 0; a < 3; ++a)
            for (int bb = 0; bb < 3; ++bb) {
                D[a * W + bb] += sc2 * M[(3 + a) * 6 + (3 + bb)];
                C[a * W + bb] += sc2 * M[(3 + a) * 6 + bb];
            }
        c[1] -= sc2 * Dz; c[2] -= sc2 * Dz;
    }
    if (k + 1 < N) { // previous-previous part of stage k+1's jerk block lands on this stage
        double M[36], sc2, Dz;
        fb_jerk_block(p, m, S, k + 1, M, sc2, Dz);
        for (int a = 0; a < 3; ++a)
            for (int bb = 0; bb < 3; ++bb) D[a * W + bb] += sc2 * M[a * 6 + bb];
        c[1] += sc2 * Dz; c[2] += sc2 * Dz;
    }
    // slacks (:211-215)
    c[3] += Wt[3];
    c[4] += 1e2 * Wt[4];
    D[4 * W + 4] += 2 * Wt[4];
    c[5] += Wt[5];
    c[6] += Wt[6];
}

// ---- AB / BL / TV stage cost; w = [v a xi...] ----
static __device__ void ab_stage_cost(const eepacc_params& p, double m, BuildSmem& S, int k)
{
    const int W = 6;
    const double* Wt = p.W_AB;
    const double w_FC = Wt[0], w_a = Wt[1], w_j = Wt[2], T = p.Ts;
    double* D = S.Dk[k]; double* C = S.Ck[k]; double* c = S.cs[k];
    for (int i = 0; i < W * W; ++i) { D[i] = 0; C[i] = 0; }
    for (int i = 0; i < W; ++i) c[i] = 0;
    if (w_FC != 0.0) {          // ABO CreateQP_AB.m:161-166
        D[0] += w_FC * 2 * p.p01 * p.F2;
        c[0] += w_FC * p.p10;
        c[1] += w_FC * p.p01 * p.lambda * m;
    }
    D[1 * W + 1] += 2 * w_a;    // CreateQP_AB.m:151-153
    if (k == 0) {               // :156-159
        D[1 * W + 1] += 2 * w_j / (T * T);
        c[1] -= 2 * w_j / T * S.a_minus1;
    } else {                    // :161-162
        D[1 * W + 1] += 2 * w_j / (T * T) * 1.0;
        C[1 * W + 1] += 2 * w_j / (T * T) * -1.0;
    }
    if (k + 1 < N) D[1 * W + 1] += 2 * w_j / (T * T) * 1.0;
    c[2] += Wt[3];              // :166-170
    c[3] += 1e2 * Wt[4];
    D[3 * W + 3] += 2 * Wt[4];
    c[4] += Wt[5];
    c[5] += Wt[6];
}

static __device__ void bl_stage_cost(const eepacc_params& p, int tv, BuildSmem& S, int k)
{
    const int W = 3;
    const double* Wt = tv ? p.W_TV : p.W_BL;
    const double w_v = Wt[0], w_a = Wt[1], w_j = Wt[2], w_f = Wt[3], Ts = p.Ts;
    double* D = S.Dk[k]; double* C = S.Ck[k]; double* c = S.cs[k];
    for (int i = 0; i < W * W; ++i) { D[i] = 0; C[i] = 0; }
    for (int i = 0; i < W; ++i) c[i] = 0;
    c[0] -= w_v;                // CreateQP_BL.m:131
    D[1 * W + 1] += 2 * w_a;    // :134-135
    if (k == 0) {               // :138-141
        D[1 * W + 1] += 2 * w_j / (Ts * Ts);
        c[1] -= 2 * w_j / Ts * S.a_minus1;
    } else {                    // :143-144
        D[1 * W + 1] += 2 * w_j / (Ts * Ts);
        C[1 * W + 1] += -2 * w_j / (Ts * Ts);
    }
    if (k + 1 < N) D[1 * W + 1] += 2 * w_j / (Ts * Ts);
    c[2] += w_f;                // :148
}

// One vehicle, one MPC step, NTB threads (the whole CTA).  Starts with a barrier; the outputs are in global
// memory after the last statement of every thread (the caller synchronises before consuming them).
template <int NTB>
__device__ void k2_vehicle(const eepacc_params& p, const eep_build_io& io, const long long veh, BuildSmem& S)
{
    const int tid = threadIdx.x;
    const int ctrl = io.ctrl;
    const int nu = (ctrl == EEPACC_FB) ? 6 : (ctrl == EEPACC_AB) ? 5 : 2;
    const int nb = (ctrl == EEPACC_FB) ? 2 : 1;
    const int W = 1 + nu;
    const int n = N * nu, m = io.m;

    const double T = p.Ts;
    __syncthreads();
    // K5: the route preview (1 KB, contiguous) is pulled into shared memory ONCE per vehicle and step with
    // coalesced 16-byte loads; the plant step (InterpPWA inside every RK4 stage) and the 20 per-stage bound
    // look-ups then walk shared memory instead of chasing the block through global memory.
    {
        static_assert(sizeof(eepacc_route) % 16 == 0, "route block must be a whole number of 16-byte words");
        const double2* src = reinterpret_cast<const double2*>(io.routes + veh * io.route_stride);
        double2* dst = reinterpret_cast<double2*>(&S.route);
        for (int e = tid; e < (int)(sizeof(eepacc_route) / 16); e += NTB) dst[e] = __ldg(src + e);
    }
    __syncthreads();
    const eepacc_route* r = &S.route;
    const double mass = p.m * r->mass_scale;
    // ---------- measurement + estimators (one thread; RunOpt_FBMPC.m:166-208) ----------
    if (tid == 0) {
        double s_meas, v_meas, a_m1, s_tv = 0, v_tv = 0, a_tv = 0, t0;
        if (io.mode == 0) {
            t0 = io.t0;
            if (io.kk == 0) {
                s_meas = p.s_init; v_meas = p.v_init; a_m1 = p.a_minus1_init;
                v_tv = 0.0; a_tv = 0.0;
            } else {
                double s = io.s_prev[veh], v = io.v_prev[veh];
                const double vp = v;
                eep_plant_step(p, r, s, v, io.Fm_prev[veh], io.Fb_prev[veh]);
                s_meas = s; v_meas = v;
                a_m1 = (v_meas - vp) / T;
                double v_tv_prev = io.v_tv_meas[veh];
                v_tv = (ctrl == EEPACC_TV) ? 0.0 : io.lead[2 * veh + 1];
                a_tv = (v_tv - v_tv_prev) / T;
            }
            if (ctrl != EEPACC_TV) s_tv = io.lead[2 * veh];
            io.v_tv_meas[veh] = v_tv;
            io.a_minus1[veh] = a_m1;
            io.x0[2 * veh] = s_meas; io.x0[2 * veh + 1] = v_meas;
        } else {
            s_meas = io.in_x0[2 * veh]; v_meas = io.in_x0[2 * veh + 1];
            a_m1 = io.in_am1[veh];
            s_tv = io.in_lead[3 * veh]; v_tv = io.in_lead[3 * veh + 1]; a_tv = io.in_lead[3 * veh + 2];
            t0 = io.in_t0[veh];
        }
        S.s_meas = s_meas; S.v_meas = v_meas; S.a_minus1 = a_m1; S.t0 = t0;
        S.s_tv = s_tv; S.v_tv = v_tv; S.a_tv = a_tv;
        const double* sp = io.prev_sol ? io.prev_sol + veh * 2 * (N + 1) : nullptr;
        int est = ctrl == EEPACC_BL ? 3 : ctrl == EEPACC_TV ? 2 : 0;
        eep_est_traj(p, est, s_meas, v_meas, a_m1, sp, sp ? sp + (N + 1) : nullptr, S.s_est, S.v_est);
        if (ctrl != EEPACC_TV) eep_est_traj(p, 1, s_tv, v_tv, a_tv, nullptr, nullptr, S.s_tv_est, S.v_tv_est);
        else for (int i = 0; i <= N; ++i) S.s_tv_est[i] = 0.0;
    }
    __syncthreads();
    // ---------- per-stage bounds (EstimateRouteAndComfortBounds) ----------
    for (int k = tid; k < N; k += NTB) {        // one stage per thread (a serial loop in the host emulation)
        int mpctype = ctrl == EEPACC_BL ? 1 : ctrl == EEPACC_TV ? 2 : 0;
        double ve = (ctrl == EEPACC_TV) ? 0.0 : S.v_est[k];         // CreateQP_TV.m:44
        eep_bounds_stage(p, r, mpctype, k, S.s_est[k], ve, S.t0, S.bnd[k]);
        double sn, cn; sincos(S.bnd[k].slope, &sn, &cn);
        S.zrg[k] = mass * p.g * (p.c_r * cn + sn);
    }
    __syncthreads();
    // ---------- dynamics A_k, D_k (FB: RunOpt_FBMPC.m:78-90,247-259 incl. the A(k) quirk) ----------
    if (tid == 0) {
        if (ctrl == EEPACC_FB) {
            double* fa = io.fbA + veh * N; double* fd = io.fbD + veh * N;
            if (io.mode == 0 && io.kk == 0) {
                for (int k = 0; k < N; ++k) {
                    fa[k] = 1 - 2 * T * p.zeta_a * p.v_init * p.v_init / p.lambda / mass;
                    fd[k] = T / p.lambda / mass * (p.zeta_a * p.v_init * p.v_init);
                }
            }
            if (p.fb_ad_quirk) {
                int row = io.kk;                     // A(k,..): k = SIMULATION step (1-based) -> kk
                if (row < N) {
                    int i = N - 1;
                    double sn, cn; sincos(S.bnd[i].slope, &sn, &cn);
                    fa[row] = 1 - 2 * T * p.zeta_a * S.v_est[i] / p.lambda / mass;
                    fd[row] = T / p.lambda / mass * (p.zeta_a * S.v_est[i] * S.v_est[i] - mass * p.g * (p.c_r * cn + sn));
                }
            } else {
                for (int i = 0; i < N; ++i) {
                    double sn, cn; sincos(S.bnd[i].slope, &sn, &cn);
                    fa[i] = 1 - 2 * T * p.zeta_a * S.v_est[i] / p.lambda / mass;
                    fd[i] = T / p.lambda / mass * (p.zeta_a * S.v_est[i] * S.v_est[i] - mass * p.g * (p.c_r * cn + sn));
                }
            }
            for (int k = 0; k < N; ++k) { S.alpha[k] = fa[k]; S.Dv[k] = fd[k]; }
        } else {
            for (int k = 0; k < N; ++k) { S.alpha[k] = 1.0; S.Dv[k] = 0.0; }
        }
        // free response d (TransformToDenseFormulation.m:55-68)
        double x0 = S.s_meas, x1 = S.v_meas;
        S.ds[0] = x0; S.dv[0] = x1;
        for (int k = 1; k <= N; ++k) {
            double n0 = x0 + T * x1;
            double n1 = S.alpha[k - 1] * x1 + S.Dv[k - 1];
            x0 = n0; x1 = n1;
            S.ds[k] = x0; S.dv[k] = x1;
        }
    }
    __syncthreads();
    // ---------- sensitivities PS/PV (TransformToDenseFormulation.m:46-52) ----------
    for (int i = tid; i < N; i += NTB) {
        double M0 = (ctrl == EEPACC_FB) ? 0.0 : .5 * T * T;           // B_i (RunOpt_FBMPC.m:89 / RunOpt_ABMPC.m:78-79)
        double M1 = (ctrl == EEPACC_FB) ? T / p.lambda / mass : T;
        for (int k = 0; k <= N; ++k) {
            if (k <= i) { S.PS[k][i] = 0.0; S.PV[k][i] = 0.0; continue; }
            S.PS[k][i] = M0; S.PV[k][i] = M1;
            if (k < N) {
                double n0 = M0 + T * M1, n1 = S.alpha[k] * M1;
                M0 = n0; M1 = n1;
            }
        }
        // stage cost blocks
        if (ctrl == EEPACC_FB) fb_stage_cost(p, mass, S, i);
        else if (ctrl == EEPACC_AB) ab_stage_cost(p, mass, S, i);
        else bl_stage_cost(p, ctrl == EEPACC_TV, S, i);
    }
    if (tid == 0) S.cs_vN = (ctrl == EEPACC_BL || ctrl == EEPACC_TV) ? -((ctrl == EEPACC_TV) ? p.W_TV[0] : p.W_BL[0]) : 0.0;
    __syncthreads();
    // ---------- t(v_k) and MV[k][j] = (Hs Psi)(v_k, j) ----------
    for (int k = tid; k <= N; k += NTB) {
        double t = 0.0;
        if (k < N) {
            t = S.Dk[k][0] * S.dv[k] + S.cs[k][0];
            if (k > 0) t += S.Ck[k][0] * S.dv[k - 1];
            if (k + 1 < N) t += S.Ck[k + 1][0] * S.dv[k + 1];
        } else t = S.cs_vN;
        S.tv[k] = t;
    }
    for (int idx = tid; idx < N * n; idx += NTB) {
        const int k = idx / n, j = idx % n;
        const int i = j / nu, cj = j % nu;
        const bool drv = cj < nb;
        double v = 0.0;
        if (drv) {
            v += S.Dk[k][0] * S.PV[k][i];
            if (k > 0) v += S.Ck[k][0] * S.PV[k - 1][i];
            if (k + 1 < N) v += S.Ck[k + 1][0] * S.PV[k + 1][i];
        }
        if (k == i) v += S.Dk[k][1 + cj];
        if (k - 1 == i) v += S.Ck[k][1 + cj];
        if (k + 1 == i && k + 1 < N) v += S.Ck[k + 1][(1 + cj) * W];
        S.MV[k][j] = v;
    }
    __syncthreads();
    // ---------- H (n x n, column-major; only when the dense conic operands are asked for) ----------
    auto Hentry = [&](int a, int j) -> double {
        const int ka = a / nu, ca = a % nu;
        const int i = j / nu, cj = j % nu;
        const bool drvj = cj < nb;
        const int r = 1 + ca;
        double v = 0.0;
        // (Hs Psi)(ctrl_a, j)
        if (drvj) {
            v += S.Dk[ka][r * W] * S.PV[ka][i];
            if (ka > 0) v += S.Ck[ka][r * W] * S.PV[ka - 1][i];
            if (ka + 1 < N) v += S.Ck[ka + 1][r] * S.PV[ka + 1][i];
        }
        if (ka == i) v += S.Dk[ka][r * W + 1 + cj];
        if (ka - 1 == i) v += S.Ck[ka][r * W + 1 + cj];
        if (ka + 1 == i) v += S.Ck[ka + 1][(1 + cj) * W + r];
        if (ca < nb) {
            double acc = 0.0;
            for (int k = ka + 1; k < N; ++k) acc += S.PV[k][ka] * S.MV[k][j];
            v += acc;
        }
        return v;
    };
    if (io.H) {
        double* Hout = io.H + veh * (long long)n * n;
        for (int idx = tid; idx < n * n; idx += NTB) Hout[idx] = Hentry(idx % n, idx / n);   // row fastest: coalesced
    }
    // ---------- structured block for the matrix-free solver path ----------
    if (io.sblk) {
        double* sb = io.sblk + veh * (long long)EEP_SBLK;
        for (int e = tid; e < (N + 1) * N; e += NTB) {
            sb[EEP_SB_PS + e] = S.PS[e / N][e % N];
            sb[EEP_SB_PV + e] = S.PV[e / N][e % N];
        }
        const int nc = 1 + nb;                          // core = [v, driven controls]
        for (int e = tid; e < N * 9; e += NTB) {
            const int k = e / 9, a = (e % 9) / 3, b = e % 3;
            const bool in = a < nc && b < nc;
            sb[EEP_SB_DC + e] = in ? S.Dk[k][a * W + b] : 0.0;
            sb[EEP_SB_CC + e] = in ? S.Ck[k][a * W + b] : 0.0;
        }
        for (int e = tid; e < N * 4; e += NTB) {
            const int k = e / 4, c = nb + e % 4;        // non-driven control c: diagonal curvature only
            sb[EEP_SB_DG + e] = (c < nu) ? S.Dk[k][(1 + c) * W + (1 + c)] : 0.0;
        }
        for (int k = tid; k < N; k += NTB) sb[EEP_SB_VEST + k] = (ctrl == EEPACC_TV) ? 0.0 : S.v_est[k];
        // scale of H: max |diag(H)| and max |Hs| (zero Hessian <=> LP)
        double hd = 0.0, hm = 0.0;
        for (int a = tid; a < n; a += NTB) hd = fmax(hd, fabs(Hentry(a, a)));
        for (int e = tid; e < N * W * W; e += NTB) {
            const int k = e / (W * W), q = e % (W * W);
            hm = fmax(hm, fmax(fabs(S.Dk[k][q]), fabs(S.Ck[k][q])));
        }
        __shared__ double red_hd[NTB / 32 > 0 ? NTB / 32 : 1], red_hm[NTB / 32 > 0 ? NTB / 32 : 1];
        for (int o = 16; o > 0; o >>= 1) { hd = fmax(hd, __shfl_xor_sync(0xffffffffu, hd, o)); hm = fmax(hm, __shfl_xor_sync(0xffffffffu, hm, o)); }
        if ((tid & 31) == 0) { red_hd[tid >> 5] = hd; red_hm[tid >> 5] = hm; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < NTB / 32; ++w) { hd = fmax(hd, red_hd[w]); hm = fmax(hm, red_hm[w]); }
            sb[EEP_SB_HSCALE] = hd; sb[EEP_SB_HMAX] = hm; sb[EEP_SB_MASS] = mass;
        }
    }
    double* gout = io.g + veh * (long long)n;
    for (int a = tid; a < n; a += NTB) {
        const int ka = a / nu, ca = a % nu, r = 1 + ca;
        double v = S.Dk[ka][r * W] * S.dv[ka] + S.cs[ka][r];
        if (ka > 0) v += S.Ck[ka][r * W] * S.dv[ka - 1];
        if (ka + 1 < N) v += S.Ck[ka + 1][r] * S.dv[ka + 1];
        if (ca < nb) {
            double acc = 0.0;
            for (int k = ka + 1; k <= N; ++k) acc += S.PV[k][ka] * S.tv[k];
            v += acc;
        }
        gout[a] = v;
    }
    if (io.qz) {    // un-condensed gradient for the stage-structured solver (k6_ocp.cuh)
        double* qz = io.qz + veh * (long long)(N + 1 + n);
        for (int k = tid; k <= N; k += NTB) qz[k] = S.tv[k];
        for (int a = tid; a < n; a += NTB) {
            const int ka = a / nu, r = 1 + a % nu;
            double v = S.Dk[ka][r * W] * S.dv[ka] + S.cs[ka][r];
            if (ka > 0) v += S.Ck[ka][r * W] * S.dv[ka - 1];
            if (ka + 1 < N) v += S.Ck[ka + 1][r] * S.dv[ka + 1];
            qz[N + 1 + a] = v;
        }
    }
    // ---------- A (structural non-zeros only) and shifted bounds ----------
    double* Aout = io.A ? io.A + veh * (long long)m * n : nullptr;
    double* lbo = io.lba + veh * (long long)m;
    double* ubo = io.uba + veh * (long long)m;
    for (int c = tid; c < m; c += NTB) {
        RowDesc rd;
        eep_row<true>(ctrl, p, mass, S, c, rd);
        double gd = 0.0;
        if (rd.kst >= 0) {
            const int k = rd.kst;
            gd = rd.cS * S.ds[k] + rd.cV * S.dv[k];
            if (Aout) for (int i = 0; i < k && i < N; ++i) {
                double v = rd.cS * S.PS[k][i] + rd.cV * S.PV[k][i];
                for (int cj = 0; cj < nb; ++cj) Aout[c + (long long)(i * nu + cj) * m] = v;
            }
        }
        if (Aout) {
            if (rd.nctl > 0) Aout[c + (long long)rd.cj0 * m] = rd.cv0;
            if (rd.nctl > 1) Aout[c + (long long)rd.cj1 * m] = rd.cv1;
            if (rd.nctl > 2) Aout[c + (long long)rd.cj2 * m] = rd.cv2;
            if (rd.kst == -2) {      // jerk rows: -1 on the driven controls of the previous stage
                const int kprev = rd.cj0 / nu - 1;
                for (int cj = 0; cj < nb; ++cj) Aout[c + (long long)(kprev * nu + cj) * m] = -1.0;
            }
        }
        lbo[c] = rd.lb - gd;
        ubo[c] = rd.ub - gd;
    }
    // dynamics for the post kernel's state roll-out
    if (io.dyn) for (int k = tid; k < N; k += NTB) { io.dyn[veh * 2 * N + k] = S.alpha[k]; io.dyn[veh * 2 * N + N + k] = S.Dv[k]; }
}
