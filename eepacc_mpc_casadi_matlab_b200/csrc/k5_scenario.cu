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


// =====================================================================================================================
// K5r -- route previews of the scenario batch built in HBM: one thread per scenario.
// Restates GenerateUseCase.m:50-107 (speed-limit staircase, curvature and slope sampled every metre) and
// SimplifyPWA.m:14-49 (drop collinear samples, then the duplicate-abscissa pass with its `doms(j)` indexing quirk,
// :43-44) operation by operation: tests bit-compare the blocks with usecase.generate_use_case.  The metre-by-metre
// sample arrays (thousands of points per route) are never stored: a sample is evaluated from the <= 12 curve / slope
// segments on the fly (the LAST segment that covers a position wins, as the reference's overwriting loop does) and
// SimplifyPWA's first pass only needs a three-sample window.  Bytes: 1.1 KB spec in, 1.6 KB route block out per scenario.
// =====================================================================================================================
#define K5R_CAP 64      // change points a pass-1 list can hold (<= 4 per segment + the two ends)

// SimplifyPWA.m second pass (:29-49) on a short list, in place; returns the new length.  doms(j) quirk kept.
__device__ int k5r_dedup(double* d, double* v, int len)
{
    double dn[K5R_CAP + 1], vn[K5R_CAP + 1];
    int cnt = 0, j = 0;
    for (int i = 0; i < len - 1; ++i) {
        if (d[i] == d[i + 1]) {
            if (v[i] == v[i + 1]) { }
            else { dn[cnt] = d[i] - .1; vn[cnt] = v[i]; ++cnt; ++j; }
        } else { dn[cnt] = d[j]; vn[cnt] = v[j]; ++cnt; ++j; }        // sic: doms(j), not doms(i)
    }
    dn[cnt] = d[len - 1]; vn[cnt] = v[len - 1]; ++cnt;
    for (int i = 0; i < cnt; ++i) { d[i] = dn[i]; v[i] = vn[i]; }
    return cnt;
}

// sample i (1-based metre) of a segment list: value of the last covering segment, else 0
__device__ __forceinline__ double k5r_sample(const double (*seg)[3], int nseg, int i, bool absval)
{
    double val = 0.0;
    for (int c = 0; c < nseg; ++c)
        if (i >= (int)seg[c][1] && i <= (int)seg[c][2]) val = absval ? fabs(seg[c][0]) : seg[c][0];
    return val;
}

// SimplifyPWA of the unit-spaced samples 1..n of a segment list; floor1e6: curvature samples below 1e-6 become 1e-6
__device__ int k5r_simplify_sampled(const double (*seg)[3], int nseg, bool curv, double* d, double* v, int* overflow)
{
    int n = 0;
    for (int c = 0; c < nseg; ++c) { const int e = (int)seg[c][2]; if (e > n) n = e; }
    auto smp = [&](int i) { double x = k5r_sample(seg, nseg, i, curv); if (curv && fabs(x) <= 1e-6) x = 1e-6; return x; };
    int len = 0;
    double vm = smp(1);
    d[len] = 1.0; v[len] = vm; ++len;
    if (n >= 2) {
        double vc = smp(2);
        for (int i = 2; i <= n - 1; ++i) {                 // interior samples (MATLAB i = 2 .. n-1)
            const double vp = smp(i + 1);
            // slopes over unit spacing: (vc - vm) / 1 and (vp - vc) / 1
            if ((vc - vm) / ((double)i - (double)(i - 1)) != (vp - vc) / ((double)(i + 1) - (double)i)) {
                if (len >= K5R_CAP) { *overflow = 1; return len; }
                d[len] = (double)i; v[len] = vc; ++len;
            }
            vm = vc; vc = vp;
        }
        if (len >= K5R_CAP) { *overflow = 1; return len; }
        d[len] = (double)n; v[len] = vc; ++len;
    }
    return k5r_dedup(d, v, len);
}

__global__ void __launch_bounds__(128) k5_route_gen_kernel(long long B, const eepacc_route_spec* __restrict__ specs,
                                                          eepacc_route* __restrict__ routes, int* __restrict__ err)
{
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
        const eepacc_route_spec& sp = specs[b];
        eepacc_route r;
        memset(&r, 0, sizeof r);
        int bad = 0;
        double d[K5R_CAP + 1], v[K5R_CAP + 1];
        // ---- speed limits (GenerateUseCase.m:52-70): staircase with a point 1 m before every change ----
        {
            double zv[EEPACC_SPEC_MAX_ZONE + 1], zs[EEPACC_SPEC_MAX_ZONE + 1];
            int nz = sp.n_zone;
            if (nz <= 0) { nz = 2; zv[0] = 1e5; zs[0] = 0; zv[1] = 1e5; zs[1] = 1; }
            else for (int i = 0; i < nz; ++i) { zv[i] = sp.zones[i][0]; zs[i] = sp.zones[i][1]; }
            for (int i = 1; i < nz; ++i) {                 // stable insertion sort by start position
                const double a = zv[i], c = zs[i];
                int k = i - 1;
                while (k >= 0 && zs[k] > c) { zv[k + 1] = zv[k]; zs[k + 1] = zs[k]; --k; }
                zv[k + 1] = a; zs[k + 1] = c;
            }
            for (int i = 0; i < nz; ++i) zv[i] = zv[i] / 3.6;
            zv[nz] = zv[nz - 1]; zs[nz] = 1e5;
            int len = 2 * nz + 1;
            d[0] = -1.0; v[0] = zv[0];
            for (int i = 1; i < 2 * nz; i += 2) {
                const int k = (i + 1) / 2;
                d[i] = zs[k - 1]; v[i] = zv[k - 1];
                d[i + 1] = zs[k] - 1; v[i + 1] = zv[k - 1];
            }
            // SimplifyPWA first pass on the short list (general spacing: IEEE division, NaN != NaN)
            double d1[2 * EEPACC_SPEC_MAX_ZONE + 2], v1[2 * EEPACC_SPEC_MAX_ZONE + 2];
            int l1 = 0;
            d1[l1] = d[0]; v1[l1] = v[0]; ++l1;
            for (int i = 1; i < len - 1; ++i) {
                const double prev = (v[i] - v[i - 1]) / (d[i] - d[i - 1]);
                const double curr = (v[i + 1] - v[i]) / (d[i + 1] - d[i]);
                if (prev != curr) { d1[l1] = d[i]; v1[l1] = v[i]; ++l1; }
            }
            d1[l1] = d[len - 1]; v1[l1] = v[len - 1]; ++l1;
            for (int i = 0; i < l1; ++i) { d[i] = d1[i]; v[i] = v1[i]; }
            len = k5r_dedup(d, v, l1);
            if (len > EEPACC_MAX_SL) bad = 1;
            else { for (int i = 0; i < len; ++i) { r.s_speedLim[i] = d[i]; r.v_speedLim[i] = v[i]; } r.n_sl = len; }
        }
        // ---- curvature (GenerateUseCase.m:73-92): |kappa| sampled every metre, floor 1e-6, two closing points ----
        {
            double one[1][3] = { { 1.0 / 1e5, 1, 2 } };
            const bool none = sp.n_curve <= 0;
            int ov = 0;
            int len = k5r_simplify_sampled(none ? one : sp.curves, none ? 1 : sp.n_curve, true, d, v, &ov);
            if (ov || len + 2 > EEPACC_MAX_CURV) bad = 2;
            else {
                for (int i = 0; i < len; ++i) { r.s_curv[i] = d[i]; r.curvature[i] = v[i]; }
                r.s_curv[len] = d[len - 1] + 1; r.curvature[len] = 1e-6;
                r.s_curv[len + 1] = d[len - 1] + 2; r.curvature[len + 1] = 1e-6;
                r.n_curv = len + 2;
            }
        }
        // ---- slope (GenerateUseCase.m:95-107): angle sampled every metre ----
        {
            double one[1][3] = { { 0.0, 1, 2 } };
            const bool none = sp.n_slope <= 0;
            int ov = 0;
            int len = k5r_simplify_sampled(none ? one : sp.slopes, none ? 1 : sp.n_slope, false, d, v, &ov);
            if (ov || len > EEPACC_MAX_SLOPE) bad = 3;
            else { for (int i = 0; i < len; ++i) { r.s_slope[i] = d[i]; r.slope[i] = v[i]; } r.n_slope = len; }
        }
        if (sp.n_stop > EEPACC_MAX_STOP || sp.n_tl > EEPACC_MAX_TL || sp.n_stop < 0 || sp.n_tl < 0) bad = 4;
        else {
            for (int i = 0; i < sp.n_stop; ++i) r.stopLoc[i] = sp.stops[i];
            for (int i = 0; i < sp.n_tl; ++i) for (int q = 0; q < 4; ++q) r.TLLoc[i][q] = sp.tls[i][q];
            r.n_stop = sp.n_stop; r.n_tl = sp.n_tl;
        }
        r.mass_scale = sp.mass_scale;
        routes[b] = r;
        if (bad) atomicMin(err, (int)(b < 0x3fffffff ? b : 0x3fffffff) * 8 + bad);
    }
}

extern "C" int eep_launch_route_gen(long long B, const eepacc_route_spec* specs, eepacc_route* routes, int* err, cudaStream_t st)
{
    long long g = (B + 127) / 128;
    if (g > 148LL * 64) g = 148LL * 64;
    if (g < 1) g = 1;
    k5_route_gen_kernel<<<(int)g, 128, 0, st>>>(B, specs, routes, err);
    return (int)cudaGetLastError();
}
