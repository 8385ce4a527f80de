// k2_build.cu -- K5+K4+K2: per-step measurement, estimators, route/comfort bounds, QP build and
// condensing for one vehicle per CTA.  Replaces, per step (MC/RunOpt_FBMPC.m:166-264):
//   RunPlantModel  -> a_minus1 / lead bookkeeping  -> EstimateVehicleTrajectory (x2)
//   -> CreateQP_{FB,AB,BL,TV} (solverToUse==1 branches) -> FB A/D update -> TransformToDenseFormulation
//
// Structured, no dense GEMM: the sparse-ordered QP is never materialised.  With
//   w_k = [v_k ; u_k]  (the only z-entries that carry curvature),  PS/PV = d s_k,v_k / d u_i
// the stage blocks D_k (w_k x w_k) and C_k (w_k x w_{k-1}) give
//   H = Psi' Hs Psi ,  g = Psi'(Hs d + cs) ,  A = Gs Psi ,  bounds -= Gs d
// entry by entry (TransformToDenseFormulation.m:87-91).  A is written only at its structural
// non-zeros (the buffer is zero-filled once by the owner), H/g/bounds fully every step.
//
// HBM layout per vehicle: H n*n, A m*n column-major (what the conic scheme / the solver reads),
// g n, lba/uba m; closed-loop scalars as SoA arrays of length B.
#include "k2_build.cuh"

#define NT_BUILD 128

__global__ void __launch_bounds__(NT_BUILD) k2_build_kernel(eepacc_params p, eep_build_io io)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BuildSmem& S = *reinterpret_cast<BuildSmem*>(smem_raw);
    for (long long veh = blockIdx.x; veh < io.B; veh += gridDim.x) k2_vehicle<NT_BUILD>(p, io, veh, S);
}

extern "C" size_t eep_build_smem_bytes() { return sizeof(BuildSmem); }

extern "C" int eep_launch_build(const eepacc_params* p, const eep_build_io* io, cudaStream_t st)
{
    size_t sm = sizeof(BuildSmem);
    static size_t optin[EEP_MAX_DEV];                    // per device (eep_optin_smem)
    if (int e = eep_optin_smem(k2_build_kernel, sm, optin)) return e;
    long long g = io->B;
    const long long cap = 148LL * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    k2_build_kernel<<<(int)g, NT_BUILD, sm, st>>>(*p, *io);
    return (int)cudaGetLastError();
}
