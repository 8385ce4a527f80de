// K5s -- lead-vehicle traces of the synthetic scenario batch, synthesised in HBM (sm_100a).
//
// every scenario's lead trace is the base driving cycle
// (Run_DrivingCycle.m:16-46 + Main.m:88, 2 Hz samples) scaled, delayed, rippled and re-thresholded, then
// integrated to a distance trace.  The host only draws the 16 perturbation numbers per scenario; this kernel
// writes the [n_steps][B][2] trace the closed loop streams (one 16-byte (s_tv, v_tv) pair per vehicle and step,
// consecutive vehicles in consecutive addresses), so a 10^6-scenario sweep uploads 128 B instead of 14 KB per
// scenario.  Bytes: n_steps * 16 B written per vehicle, nothing re-read; the base cycle (7 KB) stays in L1/L2.
//
// Arithmetic mirrors usecase.synth_lead_host (numpy, no FMA contraction) operation by operation; sin() is
// CUDA's (<= 2 ulp), so traces agree with the host's to round-off, not bit for bit (tests: 1e-9 m).
#include <cuda_runtime.h>

#include "eepacc_kernels.h"

#define K5S_VEH 32     // vehicles per CTA = lanes of a warp: consecutive vehicles -> one 512-byte store per step
#define K5S_TCH 128    // time steps per shared-memory tile
#define K5S_NW 8       // warps per CTA

// CTA = 32 vehicles x 8 warps.  The expensive part of a sample (three sin() per step) has no dependence along
// the time axis, so the 8 warps evaluate a [128 steps][32 vehicles] tile of rippled, re-thresholded speed samples
// into shared memory in parallel; warp 0 then integrates the tile step by step (the running distance is the only
// serial quantity, kept in the reference's summation order) and stores one double2 per vehicle and step.
__global__ void __launch_bounds__(K5S_VEH * K5S_NW) k5_lead_synth_kernel(long long B, int n_steps, double ts, double tv_length,
                                                                       const double* __restrict__ base, int n_base,
                                                                       const double* __restrict__ hw, int n_hw,
                                                                       const double* __restrict__ pert, double* __restrict__ lead)
{
    __shared__ double ysm[K5S_TCH][K5S_VEH];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long b = (long long)blockIdx.x * K5S_VEH + lane;
    const bool live = b < B;
    const double* q = pert + (live ? b : B - 1) * EEPACC_N_PERT;
    const double scale = q[0], gap = q[2];
    const int delay = (int)q[1];
    double A[3], om[3], ph[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        A[r] = q[3 + 3 * r];
        om[r] = __dmul_rn(6.283185307179586, q[4 + 3 * r]);        // 2*pi*f
        ph[r] = q[5 + 3 * r];
    }
    const bool use_hw = q[12] != 0.0 && hw != nullptr;
    const double* src = use_hw ? hw : base;
    const int n_src = use_hw ? n_hw : n_base;
    const int last = (n_src < n_steps ? n_src : n_steps) - 1;       // the host pads the cycle with its last sample
    double2* out = reinterpret_cast<double2*>(lead);
    double s = gap, v = 0.0;
    if (w == 0 && live) out[b] = make_double2(s - tv_length, 0.0);  // Run_DrivingCycle.m:35-37: v(1)=0, s(1)=initDist
    for (int k0 = 1; k0 < n_steps - 1; k0 += K5S_TCH) {
        const int kn = (n_steps - 1 - k0) < K5S_TCH ? (n_steps - 1 - k0) : K5S_TCH;
        for (int j = w; j < kn; j += K5S_NW) {
            const int k = k0 + j, idx = k - delay;
            double y = idx < 0 ? 0.0 : __dmul_rn(src[idx < last ? idx : last], scale);
            const double t = __dmul_rn((double)k, ts);
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const double mask = y > 0.0 ? 1.0 : 0.0;
                const double arg = __dadd_rn(__dmul_rn(om[r], t), ph[r]);
                y = __dadd_rn(y, __dmul_rn(__dmul_rn(A[r], sin(arg)), mask));
            }
            y = fmax(y, 0.0);
            if (y < 0.1) y = 0.0;                                   // Run_DrivingCycle.m:19-24 threshold
            ysm[j][lane] = y;
        }
        __syncthreads();
        if (w == 0 && live) {
            for (int j = 0; j < kn; ++j) {
                v = ysm[j][lane];
                s = __dadd_rn(s, __dmul_rn(ts, v));                 // :40-43
                out[(long long)(k0 + j) * B + b] = make_double2(s - tv_length, v);
            }
        }
        __syncthreads();
    }
    if (w == 0 && live && n_steps > 1) out[(long long)(n_steps - 1) * B + b] = make_double2(s - tv_length, v);   // :44-46 last repeats
}

int eep_launch_lead_synth(long long B, int n_steps, double ts, double tv_length, const double* base, int n_base,
                          const double* hw, int n_hw, const double* pert, double* lead, cudaStream_t st)
{
    const long long grid = (B + K5S_VEH - 1) / K5S_VEH;
    k5_lead_synth_kernel<<<(unsigned)grid, K5S_VEH * K5S_NW, 0, st>>>(B, n_steps, ts, tv_length, base, n_base, hw, n_hw, pert, lead);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
