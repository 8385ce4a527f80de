// k6_ocp.cu -- K6 launch: the stage-structured interior-point QP solver of k6_ocp.cuh, one vehicle per thread.
//
// Opt-in alternative to K3/K3r in the closed loop (eepacc_mem_set_solver(mem, EEPACC_SOLVER_OCP)); it stands where
// the reference switches from condensing + dense qpOASES to its structured solvers (OPTsettings.solverToUse = 0 / 2,
// MC/Settings.m:112, RunOpt_FBMPC.m:96-121, RunOpt_ABMPC.m:87-117).  K2 -> K6 -> K4: same operands in, same
// (x, cost, status, iter) out as the dense kernels.
//
// Launch shape: 32 threads per CTA so that a batch of 8192 vehicles already spreads over every SM (256 warps over
// 148 SMs); larger batches use 64 / 128 threads.  The vehicles share nothing but the read-only row table (one copy
// per CTA in shared memory); there is no barrier after the set-up.
#include "k6_ocp.cuh"

template <int NU, int NB>
__global__ void __launch_bounds__(128) k6_ocp_kernel(const __grid_constant__ eepacc_params p, const eep_ocp_io io)
{
    // the packed row table and its coefficient tables are the same for every vehicle: one copy per CTA in shared memory
    __shared__ unsigned long long s_tab[EEPACC_MAX_NCON];
    __shared__ double s_rowc[96];
    for (int i = threadIdx.x; i < io.m; i += blockDim.x) s_tab[i] = io.rowtab[i];
    for (int i = threadIdx.x; i < 96; i += blockDim.x) s_rowc[i] = io.rowc[i];
    __syncthreads();
    const long long veh = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (veh >= io.B) return;
    constexpr int NH = EEP_N, n = NH * NU;
    k6::Args A;
    A.sb = io.sblk + veh * EEP_SBLK;
    A.alpha = io.dyn + veh * 2 * NH;
    A.qz = io.qz + veh * (NH + 1 + n);
    A.lba = io.lba + veh * io.m;
    A.uba = io.uba + veh * io.m;
    A.rowtab = s_tab; A.rowc = s_rowc;
    A.work = io.work + veh; A.ws = io.work_stride;
    A.m = io.m;
    A.T = p.Ts;
    A.M0 = (io.ctrl == EEPACC_FB) ? 0.0 : 0.5 * p.Ts * p.Ts;                  // B_k: RunOpt_ABMPC.m:78-79 / RunOpt_FBMPC.m:89
    A.M1 = (io.ctrl == EEPACC_FB) ? p.Ts / p.lambda / A.sb[EEP_SB_MASS] : p.Ts;
    A.mu0 = io.mu0; A.eps_mu = io.eps_mu; A.eps_p = io.eps_p; A.eps_d = io.eps_d; A.eps_accept = io.eps_accept;
    A.maxit = io.maxit; A.warm = 0;
    A.gamma = 1e-3; A.tau0 = 1.0;       // measured on the bench batch: see DESIGN.md, K6
    double cost = 0.0; int it = 0;
    const int st = k6::solve<NU, NB>(A, io.x + veh * n, cost, it);
    io.cost[veh] = cost; io.status[veh] = (double)st; io.iter[veh] = (double)it;
    if (io.counters) {
        unsigned long long* c = io.counters + veh * 4;
        c[0] += (unsigned long long)it;
        if (st) c[1] += 1ull;
    }
}

extern "C" long long eep_ocp_work_items(int ctrl, int n, int m)
{
    if (ctrl == EEPACC_FB) return k6::work_items(n, m, 6, 2);
    if (ctrl == EEPACC_AB) return k6::work_items(n, m, 5, 1);
    return 0;
}

extern "C" int eep_launch_ocp(const eepacc_params* p, const eep_ocp_io* io, cudaStream_t st)
{
    if (io->B <= 0) return 0;
    int nt = io->B <= 148LL * 32 * 4 ? 32 : io->B <= 148LL * 64 * 8 ? 64 : 128;
    const long long grid = (io->B + nt - 1) / nt;
    if (io->ctrl == EEPACC_FB) k6_ocp_kernel<6, 2><<<(unsigned)grid, nt, 0, st>>>(*p, *io);
    else if (io->ctrl == EEPACC_AB) k6_ocp_kernel<5, 1><<<(unsigned)grid, nt, 0, st>>>(*p, *io);
    else return (int)cudaErrorInvalidValue;
    return (int)cudaGetLastError();
}
