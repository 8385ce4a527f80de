// k3r_emu.cpp -- TEST INFRASTRUCTURE: single-thread host emulation of the reduced-space QP kernel.
//
// Compiles eepacc_mpc_casadi_matlab_b200/csrc/k3r_qp.cu with g++ (NT = 1: every "for (i = threadIdx.x; i < n; i += NT)"
// loop becomes a serial loop, barriers vanish, warp-shuffle code paths are replaced by their scalar
// statements inside the .cu).  It exists so that the solver LOGIC can be replayed and debugged on a
// machine without a GPU (tests/test_k3r_host_emu.py feeds it QP operands recorded from K2).  It is never
// loaded by the product package; the product path has no CPU fallback.
#define EEP_HOST_EMU 1
#include <cuda_runtime.h>
#include <cstdlib>
#include <cstring>
#include <cmath>
#ifndef __grid_constant__
#define __grid_constant__
#endif
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#define NT 1
#define EMU_ONLY(...) __VA_ARGS__
struct EmuIdx { int x; };
static const EmuIdx threadIdx = { 0 }, blockIdx = { 0 }, gridDim = { 1 };
static inline void __syncthreads() {}
#include "../../eepacc_mpc_casadi_matlab_b200/csrc/k3r_qp.cu"

extern "C" int k3r_emu_solve(const eepacc_params* p, const eep_qp_io* io_in)
{
    eep_qp_io io = *io_in;
    unsigned long long* tab = (unsigned long long*)calloc((size_t)io.m + 1, 8);
    double consts[96];
    const int rc = k3r_build_rowtab(io.ctrl, *p, io.m, tab, consts);
    if (rc) { free(tab); return 100 + rc; }
    io.rowtab = tab; io.rowc = consts;
    double* hs = (double*)calloc((size_t)io.B * 4 * (size_t)(io.n + io.m) + 1, 8);
    io.hscr = hs;
    const size_t sm = eep_qpr_smem_bytes(io.n, io.m, io.fc);
    unsigned char* buf = (unsigned char*)calloc(sm + 64, 1);
    if (!buf) { free(tab); return -1; }
    if (io.ctrl == EEPACC_FB) k3r_qp_kernel<6, 2>(io, buf);
    else if (io.ctrl == EEPACC_AB) k3r_qp_kernel<5, 1>(io, buf);
    else k3r_qp_kernel<2, 1>(io, buf);
    free(buf); free(tab); free(hs);
    return 0;
}
extern "C" long long k3r_emu_sizeof_io(void) { return (long long)sizeof(eep_qp_io); }
