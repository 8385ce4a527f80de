// k6_ocp.cu -- K6 launch: the stage-structured interior-point QP solver of k6_ocp.cuh, one vehicle per thread.
//
// Opt-in alternative to K3/K3r in the closed loop (eepacc_mem_set_solver(mem, EEPACC_SOLVER_OCP)); it stands where
// the reference switches from condensing + dense qpOASES to its structured solvers (OPTsettings.solverToUse = 0 / 2,
// MC/Settings.m:112, RunOpt_FBMPC.m:96-121, RunOpt_ABMPC.m:87-117).  K2 -> K6 -> K4: same operands in, same
// (x, cost, status, iter) out as the dense kernels.
//
// Launch shape (eep_launch_ocp): few vehicles per warp, a bounded number of vehicles in flight per SM, maximum L1.
// Nothing is shared between threads: no shared memory, no barrier.
#include "k6_ocp.cuh"
#include <cstdlib>

template <int NU, int NB>
__global__ void __launch_bounds__(128) k6_ocp_kernel(const __grid_constant__ eepacc_params p, const eep_ocp_io io)
{
    constexpr int NH = EEP_N, n = NH * NU;
    // a bounded number of vehicles is in flight (grid-stride loop): each thread keeps ~3 KB of stage blocks in local
    // memory, and what is in flight per SM has to stay inside the L1 (eep_launch_ocp)
    for (long long veh = (long long)blockIdx.x * blockDim.x + threadIdx.x; veh < io.B; veh += (long long)gridDim.x * blockDim.x) {
        k6::Args A;
        A.sb = io.sblk + veh * EEP_SBLK;
        A.alpha = io.dyn + veh * 2 * NH;
        A.qz = io.qz + veh * (NH + 1 + n);
        A.lba = io.lba + veh * io.m;
        A.uba = io.uba + veh * io.m;
        A.rowtab = io.rowtab; A.rowc = io.rowc;      // 5 KB, the same for every vehicle: one copy in the SM's L1
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
    // Measured (profiles/r2w, r2x): with 32 vehicles per warp and every vehicle of the batch in flight a solve took 0.45 s
    // -- ~300 cycles per local-memory access, i.e. the per-thread stage blocks (3 KB x 64 threads per SM and more) were
    // served by the L2, one access at a time.  So: (a) few vehicles per warp -- the solve is a chain of dependent
    // operations, lanes cost nothing, warps hide latency: K6_LANES (8) vehicles per CTA puts 7 warps on an SM at 8192
    // vehicles instead of 2; (b) at most K6_INFLIGHT (64) vehicles per SM are in flight, the rest wait in the grid-stride
    // loop, so the local memory in use stays inside the L1; (c) the L1 / shared-memory split is set to maximum L1.
    int lanes = 8, inflight = 64;
    if (const char* e = getenv("EEPACC_K6_LANES")) { int v = atoi(e); if (v >= 1 && v <= 128) lanes = v; }
    if (const char* e = getenv("EEPACC_K6_INFLIGHT")) { int v = atoi(e); if (v >= 8 && v <= 2048) inflight = v; }
    long long grid = (io->B + lanes - 1) / lanes;
    const long long cap = (148LL * inflight + lanes - 1) / lanes;
    if (grid > cap) grid = cap;
    static bool carve[EEP_MAX_DEV][2] = {};
    int dev = 0;
    const bool fb = io->ctrl == EEPACC_FB;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < EEP_MAX_DEV && !carve[dev][fb]) {
        if (fb) cudaFuncSetAttribute(k6_ocp_kernel<6, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
        else cudaFuncSetAttribute(k6_ocp_kernel<5, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
        carve[dev][fb] = true;
    }
    if (fb) k6_ocp_kernel<6, 2><<<(unsigned)grid, lanes, 0, st>>>(*p, *io);
    else if (io->ctrl == EEPACC_AB) k6_ocp_kernel<5, 1><<<(unsigned)grid, lanes, 0, st>>>(*p, *io);
    else return (int)cudaErrorInvalidValue;
    return (int)cudaGetLastError();
}
