// eepacc_kernels.h -- host-visible launch interface of the sm_100a kernels (internal to the library).
#pragma once
#include <cuda_runtime.h>
#include "../../include/eepacc.h"

#define EEPACC_MAX_NVAR 120     // FB: 6*20
#define EEPACC_MAX_NCON 522     // FB: 26*20+2
#define EEPACC_LOG_W    12      // z1[8], a, status, iter, cost
#define EEP_ACC_W       36      // running KPI accumulators per vehicle (k4_post.cuh); ONE definition for the kernels and the host side

// Structured per-vehicle QP block written by K2 for K3's matrix-free path (doubles):
//   PS/PV [(N+1) x N]  d s_k, v_k / d (driven control of stage i);  VEST [N] v_est(k);
//   DC/CC [N x 3 x 3]  stage Hessian blocks on the core [v, driven controls] (D_k) and the coupling
//                      with stage k-1 (C_k: rows stage k, cols stage k-1);  DG [N x 4] diagonal
//                      curvature of the non-driven controls;  HMAX max|Hs|, HSCALE max|diag H|, MASS.
#define EEP_SB_PS     0
#define EEP_SB_PV     420
#define EEP_SB_VEST   840
#define EEP_SB_DC     860
#define EEP_SB_CC     1040
#define EEP_SB_DG     1220
#define EEP_SB_HMAX   1300
#define EEP_SB_HSCALE 1301
#define EEP_SB_MASS   1302
#define EEP_SBLK      1304

// K2 (k2_build.cu)
struct eep_build_io {
    long long B;
    int ctrl, m, mode, kk;      // mode 0: closed loop (plant + bookkeeping inside), 1: inputs given
    double t0;
    const eepacc_route* routes; long long route_stride;
    // closed-loop state, SoA [B]
    double *s_prev, *v_prev, *Fm_prev, *Fb_prev, *a_minus1, *v_tv_meas;
    double *fbA, *fbD;          // [B][N] FB dynamics rows (A(k,2,2), D(k,2)) incl. the A(k) quirk
    const double* prev_sol;     // [B][2][N+1] previous predicted (s, v) (estimator mode 2 only)
    const double* lead;         // [B][2] this step's (s_tv, v_tv)
    // API-mode inputs
    const double *in_x0, *in_am1, *in_lead, *in_t0;
    // outputs
    double *x0;                 // [B][2] measured state
    double *H, *g, *A, *lba, *uba;   // H, A may be null (matrix-free closed loop)
    double *sblk;               // [B][EEP_SBLK] structured block (may be null)
    double *dyn;                // [B][2][N] alpha_k, Dv_k for the post kernel (may be null)
    double *qz;                 // [B][N + 1 + n] gradient of the un-condensed QP at the free response: tv_k = (Hs d + cs)(v_k),
                                // k = 0..N, then (Hs d + cs)(u) -- what the stage-structured solver K6 needs (may be null)
};

// K3 (k3_qp.cu)
struct eep_qp_io {
    long long B;
    int n, m;
    int ctrl;
    const double *H, *g, *A, *lba, *uba, *lbx, *ubx;   // lbx/ubx may be null (= +-inf); H, A null in matrix-free mode
    const double *sblk;         // [B][EEP_SBLK] structured block from K2: non-null selects the matrix-free operators
    // warm-start state (device, owned by the mem object)
    double *wx, *wy;            // [B][n]
    short *widx;                // [B][n]   constraint index per Q column
    signed char *wside;         // [B][n]
    int *wnA, *warm;            // [B]
    // outputs
    double *x, *cost, *lam_a, *lam_x, *status, *iter;   // lam_* may be null
    unsigned long long *counters;   // [B][4] working-set changes, failed, exchanges, refactorisations
    double *dbg; int dbg_cap;       // optional per-iteration trace of vehicle 0: tau, key, nA, exchanges
    // reduced-space kernel (k3r_qp.cu): capacity of the free-variable factorisation and the list of
    // vehicles it hands over to the full-space kernel; the full-space kernel, when todo_list is
    // non-null, only solves the vehicles todo_list[0 .. *todo_cnt)
    int fc;
    int *defer_list, *defer_cnt;
    const int *todo_list, *todo_cnt;
    const unsigned long long* rowtab;   // [m] packed constraint rows of the controller (eep_qpr_build_rowtab)
    const double* rowc;                 // [96] their coefficient tables
    double* hscr;                       // [B][4 (n + m)] homotopy state of the reduced-space kernel (a x, current bounds, a dx)
    int todo_grid;                      // CTAs that walk todo_list (0: 148).  These launches need a whole SM's shared
                                        // memory per CTA; a wide grid stalls the slice until that many SMs have drained
};

// K6 (k6_ocp.cu): stage-structured interior-point QP solver, one vehicle per thread (opt-in: eepacc_mem_set_solver)
struct eep_ocp_io {
    long long B;
    int n, m, ctrl;
    const double *sblk, *dyn, *qz, *lba, *uba;      // K2 outputs: [B][EEP_SBLK], [B][2][N], [B][N+1+n], [B][m], [B][m]
    const unsigned long long* rowtab;               // [m] packed constraint rows (shared with K3r)
    const double* rowc;                             // [96]
    double* work; long long work_stride;            // [eep_ocp_work_items][work_stride] per-vehicle work space, vehicle fastest
    double *x, *cost, *status, *iter;               // [B][n], [B], [B], [B]
    unsigned long long* counters;                   // [B][4]: interior-point iterations, failed (may be null)
    double mu0, eps_mu, eps_p, eps_d, eps_accept;
    int maxit;
};

// post kernel (k4_post.cu): z recovery, force allocation, logging, KPI accumulation
struct eep_post_io {
    long long B;
    int ctrl, n, kk;
    const eepacc_route* routes; long long route_stride;
    const double *x, *x0, *dyn, *status, *iter, *cost;
    double *s_prev, *v_prev, *Fm_prev, *Fb_prev;
    double *prev_sol;           // [B][2][N+1]
    double *log;                // [B][EEPACC_LOG_W] for this step (may be null)
    double *kpi_acc;            // [B][ACC_W] running KPI accumulators
    const double* lead;         // [B][2] this step's (s_tv, v_tv) for the headway KPIs (may be null: no lead)
};

// fused closed loop (k3r_qp.cu: k5_cycle_kernel): one CTA takes a vehicle through the MPC steps
// [k_begin, k_end) -- K5/K4/K2 build, reduced-space QP (hand-overs solved in place at full capacity from a
// global work area), post -- with no synchronisation between vehicles
struct eep_cycle_io {
    eep_build_io b;             // whole batch (no slice offsets); kk, t0, lead are set per step by the kernel
    eep_qp_io q;
    eep_post_io p;              // kk, log are set per step by the kernel
    const double* lead_all;     // [n_steps][B][2] or null (TVMPC)
    long long lead_stride;      // doubles per step
    double* log_all;            // [n_steps][B][EEPACC_LOG_W] or null
    long long log_stride;
    int k_begin, k_end;
    double t0_begin;            // time of step k_begin
    unsigned long long* work_counter;   // vehicles are handed out dynamically
    unsigned char* big;         // [grid][big_bytes] work area of the hand-over solves
    unsigned long long big_bytes;
    int fc_big;
    unsigned long long* phase;  // [4] clock64() cycles summed over all vehicle-steps: build, QP, hand-over QP, post
};

#ifdef __cplusplus
// Dynamic shared-memory opt-in of a kernel, cached PER DEVICE (the attribute is per device; calling
// cudaFuncSetAttribute before each of the ~55 000 launches of a cycle costs several per cent of the cycle).
// `slot` is a per-kernel static array of EEP_MAX_DEV entries, zero-initialised.
#define EEP_MAX_DEV 64
template <class K> static inline int eep_optin_smem(K kernel, size_t bytes, size_t* slot)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= EEP_MAX_DEV) dev = -1;
    if (dev >= 0 && bytes <= slot[dev]) return 0;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0) slot[dev] = bytes;
    return 0;
}
#endif

#ifdef __cplusplus
extern "C" {
#endif
int eep_launch_cycle(const eepacc_params* p, const eep_cycle_io* cy, int grid, cudaStream_t st);
int eep_cycle_max_ctas_per_sm(const eep_cycle_io* cy);
size_t eep_build_smem_bytes();
void eep_launch_f(const eepacc_params* p, long long B, const double* x, const double* u, double* xdot, double* L, cudaStream_t st);
void eep_launch_F(const eepacc_params* p, long long B, const double* x, const double* u, double* xf, double* qf, cudaStream_t st);
void eep_launch_Fjac(const eepacc_params* p, long long B, const double* x, const double* u, double* xf, double* qf, double* J, cudaStream_t st);
void eep_launch_power(const eepacc_params* p, long long B, int fifth, const double* Fm, const double* rpm, double* P, cudaStream_t st);
void eep_launch_plant(const eepacc_params* p, long long B, const eepacc_route* routes, long long route_stride,
                      const double* x, const double* u, double* xn, cudaStream_t st);
int eep_launch_build(const eepacc_params* p, const eep_build_io* io, cudaStream_t st);
int eep_launch_qp(const eepacc_params* p, const eep_qp_io* io, cudaStream_t st);
int eep_launch_qpr(const eepacc_params* p, const eep_qp_io* io, cudaStream_t st);
size_t eep_qpr_smem_bytes(int n, int m, int fc);
int eep_qpr_build_rowtab(int ctrl, const eepacc_params* p, int m, unsigned long long* tab_dev, double* consts_dev, cudaStream_t st);
size_t eep_qp_smem_bytes(int n, int m, int smode);
int eep_launch_post(const eepacc_params* p, const eep_post_io* io, cudaStream_t st);
int eep_launch_ocp(const eepacc_params* p, const eep_ocp_io* io, cudaStream_t st);
long long eep_ocp_work_items(int ctrl, int n, int m);
int eep_launch_lead_synth(long long B, int n_steps, double ts, double tv_length, const double* base, int n_base,
                          const double* hw, int n_hw, const double* pert, double* lead, cudaStream_t st);
int eep_launch_route_gen(long long B, const eepacc_route_spec* specs, eepacc_route* routes, int* err, cudaStream_t st);
int eep_launch_kpi_final(const eepacc_params* p, long long B, int n_steps, const double* kpi_acc, double* kpi, cudaStream_t st);
double eep_dfma_peak(int device);
#ifdef __cplusplus
}
#endif
