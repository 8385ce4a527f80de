// k4_post.cuh -- device code that closes one MPC step for ONE vehicle (see k4_post.cu); shared by the per-step
// kernel and the fused closed-loop kernel (k3r_qp.cu: k5_cycle_kernel).
#pragma once
#include "eepacc_dev.cuh"
#include "eepacc_kernels.h"

#ifndef N
#define N EEP_N
#endif
#define ACC_W EEP_ACC_W

__device__ __forceinline__ void k4_vehicle(const eepacc_params& p, const eep_post_io& io, const long long veh)
{
    const int ctrl = io.ctrl, n = io.n;
    const int nu = n / N;
    const int nb = (ctrl == EEPACC_FB) ? 2 : 1;
    const double T = p.Ts;
    const eepacc_route* r = io.routes + veh * io.route_stride;
    const double mass = p.m * r->mass_scale;
    const double* x = io.x + veh * (size_t)n;
    const double* alpha = io.dyn + veh * 2 * N;
    const double* Dv = alpha + N;
    const double s0 = io.x0[2 * veh], v0 = io.x0[2 * veh + 1];
    const double Bs = (ctrl == EEPACC_FB) ? 0.0 : .5 * T * T;
    const double Bv = (ctrl == EEPACC_FB) ? T / p.lambda / mass : T;
    // state roll-out = rows s_k, v_k of Psi x + d
    double* ps = io.prev_sol + veh * 2 * (N + 1);
    double* pv = ps + (N + 1);
    double s = s0, v = v0;
    ps[0] = s; pv[0] = v;
    for (int k = 0; k < N; ++k) {
        double drive = 0.0;
        for (int c = 0; c < nb; ++c) drive += x[k * nu + c];
        double sn = s + T * v + Bs * drive;
        double vn = alpha[k] * v + Bv * drive + Dv[k];
        s = sn; v = vn;
        ps[k + 1] = s; pv[k + 1] = v;
    }
    double Fm, Fb, a_out, xi0 = 0, xi1 = 0, xi2 = 0, xi3 = 0;
    if (ctrl == EEPACC_FB) {
        Fm = x[0]; Fb = x[1]; xi0 = x[2]; xi1 = x[3]; xi2 = x[4]; xi3 = x[5];
        a_out = (io.kk > 0) ? (v0 - io.v_prev[veh]) / T : 0.0;
    } else {
        if (ctrl == EEPACC_AB) { xi0 = x[1]; xi1 = x[2]; xi2 = x[3]; xi3 = x[4]; }
        else xi3 = x[1];
        eep_force_alloc(p, r, s0, v0, x[0], Fm, Fb, a_out);
    }
    io.s_prev[veh] = s0; io.v_prev[veh] = v0; io.Fm_prev[veh] = Fm; io.Fb_prev[veh] = Fb;
    const double st = io.status[veh];
    if (io.log) {
        double* o = io.log + veh * EEPACC_LOG_W;
        o[0] = s0; o[1] = v0; o[2] = Fm; o[3] = Fb; o[4] = xi0; o[5] = xi1; o[6] = xi2; o[7] = xi3;
        o[8] = a_out; o[9] = st; o[10] = io.iter[veh]; o[11] = io.cost[veh];
    }
    // running KPI accumulators
    double* A = io.kpi_acc + veh * ACC_W;
    if (io.kk == 0) for (int i = 0; i < ACC_W; ++i) A[i] = 0.0;
    const int i1 = io.kk + 1;                       // 1-based step index
    if (A[4] == 0.0 && i1 >= 2 && A[3] < p.cutOffDist && s0 > p.cutOffDist) {
        A[4] = 1.0; A[5] = (double)(i1 - 1);
        A[13] = A[1];                               // E_opt(distInd-1)
        A[14] = A[6]; A[15] = A[7]; A[16] = A[8];   // a statistics over 1..distInd
    }
    A[9] = A[6]; A[10] = A[7]; A[11] = A[8];        // statistics through the previous step
    A[2] = A[1]; A[1] = A[0];
    const double rpm = (30.0 / EEP_PI) * v0 * p.phi;
    A[0] += T * eep_power5(p.b_fifth, Fm, rpm);
    A[3] = s0;
    if (i1 == 1) { A[6] = a_out; A[7] = a_out; }
    else { A[6] = fmax(A[6], a_out); A[7] = fmin(A[7], a_out); }
    A[8] += a_out * a_out;
    if (st != 0.0) A[12] += 1.0;
    // jerk j(i) = (a(i+1) - a(i)) / Ts (RunOpt_FBMPC.m:339), statistics over j(1..distInd) (Main.m:247-250):
    // [20] a of the previous step, [21..23] running max / min / sum of squares, [24..26] their values at the cut-off
    if (i1 >= 2) {
        const double j = (a_out - A[20]) / T;
        if (i1 == 2) { A[21] = j; A[22] = j; }
        else { A[21] = fmax(A[21], j); A[22] = fmin(A[22], j); }
        A[23] += j * j;
    }
    A[20] = a_out;
    if (A[4] == 1.0) { A[4] = 2.0; A[24] = A[21]; A[25] = A[22]; A[26] = A[23]; }   // cut-off found at THIS step: j(distInd) is the newest entry
    // constraint violations: steps with an active soft constraint ([27..30]) and headway against the lead ([31..33])
    if (xi0 > 1e-6) A[27] += 1.0;
    if (xi1 > 1e-6) A[28] += 1.0;
    if (xi2 > 1e-6) A[29] += 1.0;
    if (xi3 > 1e-6) A[30] += 1.0;
    if (io.kk == 0) A[33] = INFINITY;
    if (io.lead) {
        const double gap = io.lead[2 * veh] - s0;
        if (isfinite(gap)) {
            if (gap < p.h_min) A[31] += 1.0;
            if (gap < p.tau_min * v0) A[32] += 1.0;
            A[33] = fmin(A[33], gap);
        }
    }
}
