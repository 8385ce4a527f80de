// k4_post.cu -- closes one MPC step: z recovery, force allocation, logging and running KPIs.
//   z = Psi x + d and first-stage extraction        MC/RunOpt_FBMPC.m:287-313
//   AB/BL/TV force allocation + realised accel      MC/RunOpt_ABMPC.m:288-323
//   KPI block                                       MC/Main.m:136-148,218-258, RunOpt_FBMPC.m:334-339
// One vehicle per thread; SoA state arrays are read/written fully coalesced.
#include "eepacc_dev.cuh"
#include "eepacc_kernels.h"

#define N EEP_N
#define ACC_W 20

__global__ void __launch_bounds__(128) k4_post_kernel(eepacc_params p, eep_post_io io)
{
    const int ctrl = io.ctrl, n = io.n;
    const int nu = n / N;
    const int nb = (ctrl == EEPACC_FB) ? 2 : 1;
    const double T = p.Ts;
    for (long long veh = (long long)blockIdx.x * blockDim.x + threadIdx.x; veh < io.B;
         veh += (long long)gridDim.x * blockDim.x) {
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
    }
}

__global__ void __launch_bounds__(128) k4_kpi_final_kernel(eepacc_params p, long long B, int n_steps,
                                                          const double* __restrict__ acc, double* __restrict__ kpi)
{
    for (long long veh = (long long)blockIdx.x * blockDim.x + threadIdx.x; veh < B;
         veh += (long long)gridDim.x * blockDim.x) {
        const double* A = acc + veh * ACC_W;
        double* K = kpi + veh * EEPACC_N_KPI;
        double distInd, Ecut, amax, amin, a2;
        if (A[4] != 0.0) { distInd = A[5]; Ecut = A[13]; amax = A[14]; amin = A[15]; a2 = A[16]; }
        else { distInd = (double)(n_steps - 1); Ecut = A[2]; amax = A[9]; amin = A[10]; a2 = A[11]; }
        if (distInd < 1) distInd = 1;
        K[0] = A[0]; K[1] = A[3];
        K[2] = 0.1 * floor(distInd * p.Ts * 10 + 0.5);
        K[3] = Ecut; K[4] = amax; K[5] = amin; K[6] = sqrt(a2 / distInd); K[7] = A[12];
    }
}

extern "C" int eep_launch_post(const eepacc_params* p, const eep_post_io* io, cudaStream_t st)
{
    long long g = (io->B + 127) / 128;
    if (g > 148 * 16) g = 148 * 16;
    if (g < 1) g = 1;
    k4_post_kernel<<<(int)g, 128, 0, st>>>(*p, *io);
    return (int)cudaGetLastError();
}

extern "C" int eep_launch_kpi_final(const eepacc_params* p, long long B, int n_steps, const double* kpi_acc,
                                    double* kpi, cudaStream_t st)
{
    long long g = (B + 127) / 128;
    if (g > 148 * 16) g = 148 * 16;
    if (g < 1) g = 1;
    k4_kpi_final_kernel<<<(int)g, 128, 0, st>>>(*p, B, n_steps, kpi_acc, kpi);
    return (int)cudaGetLastError();
}

// ---- FP64 FMA peak microbenchmark: register-resident dependent-free DFMA chains ----
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters)
{
    double a0 = threadIdx.x * 1e-9 + 1.0, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 0.999999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

extern "C" double eep_dfma_peak(int device)
{
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return -1.0;
    const int blocks = 148 * 8, threads = 256, iters = 1 << 16;
    double* d = nullptr;
    if (cudaMalloc(&d, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    dfma_peak_kernel<<<blocks, threads>>>(d, 1024);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    dfma_peak_kernel<<<blocks, threads>>>(d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaFree(d);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (ms <= 0) return -1.0;
    double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
    return flops / (ms * 1e-3) / 1e9;
}
