// k6_ocp.cu -- K6 launch: the stage-structured interior-point QP solver of k6_ocp.cuh, one vehicle per thread.
//
// Opt-in alternative to K3/K3r in the closed loop (eepacc_mem_set_solver(mem, EEPACC_SOLVER_OCP)); it stands where
// the reference switches from condensing + dense qpOASES to its structured solvers (OPTsettings.solverToUse = 0 / 2,
// MC/Settings.m:112, RunOpt_FBMPC.m:96-121, RunOpt_ABMPC.m:87-117).  K2 -> K6 -> K4: same operands in, same
// (x, cost, status, iter) out as the dense kernels.
//
// Launch shape (eep_launch_ocp): 32 vehicles per warp, measured; maximum L1.
// Nothing is shared between threads: no shared memory, no barrier.
#include "k6_ocp.cuh"
#include <cstdlib>

template <int NU, int NB>
__global__ void __launch_bounds__(128) k6_ocp_kernel(const __grid_constant__ eepacc_params p, const eep_ocp_io io)
{
    constexpr int NH = EEP_N, n = NH * NU;
    // grid-stride loop (the number of vehicles in flight is a launch knob: eep_launch_ocp)
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
    // Launch shape, measured on a B200 (profiles/r2y2_k6_sweep.txt, FBMPC, 8192 vehicles, ms per MPC step):
    //   vehicles per warp   32: 473    16: 490    8: 584    4: 886    2: 1538      (vehicles in flight per SM: no effect)
    // One warp needs ~0.9e9 cycles for its 32 solves however few warps share the SM, and spreading the vehicles over
    // more warps only adds warp-instructions: the kernel is bound by the chain of dependent local- / global-memory
    // accesses of ONE thread (stage blocks in local memory, row state in the [item][B] work space), ~11 M instructions at
    // ~80 cycles each.  A fast version needs the stage blocks in registers / shared memory and the rows of a vehicle
    // spread over the lanes of a warp; this one keeps 32 vehicles per warp and the grid-stride loop.
    int lanes = 32, inflight = 2048;
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
