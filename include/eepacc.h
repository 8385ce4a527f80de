/*
 * eepacc.h -- C ABI of the B200-native batched EEPACC MPC engine.
 *
 * Every entry point keeps the calling convention of a CasADi-generated function,
 *
 *     int F(const double** arg, double** res, long long* iw, double* w, int mem);
 *
 * i.e. the convention that the reference evaluates through
 * casadi_c_eval_id(id,arg,res,iw,w,mem)     (ACCMPC-ABO_CasADi/casadi_fun.c:170,
 * casadi_3_6_3/include/casadi/casadi_c.h:128-129), with the function id folded into
 * the symbol name and ONE extension: a leading batch dimension.
 *
 *   - arg[i] / res[i] point to `double[B][nnz_i]` (batch-leading; the slice of one
 *     vehicle is exactly what one reference call would see, dense column-major).
 *   - iw[0] = B (batch), iw[1] = flags (EEPACC_FLAG_*).  F_work() reports sz_iw >= 2.
 *   - `mem` is the handle returned by F_checkout() (casadi_c.h:51-64 semantics: one
 *     per host thread / CUDA stream).  It owns the device-resident per-vehicle state
 *     (plant state, warm-start working set, FB A/D quirk rows) and the scenario data.
 *   - return 0 = success, non-zero = failure (casadi_fun.c:64,170-172).  Per-vehicle
 *     solver status is an OUTPUT (res[...]), never the return code
 *     (error_on_fail=false, RunOpt_FBMPC.m:127,281).
 *
 * There is no CPU fallback: every F() launches sm_100a kernels and fails with
 * EEPACC_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef EEPACC_H
#define EEPACC_H

#ifdef __cplusplus
extern "C" {
#endif

#define EEPACC_N_HOR        20   /* Settings.m:99-100 (sum(moveBlockingSettings)); only horizon with golden data */
#define EEPACC_MAX_SL       16   /* speed-limit PWA points  (GenerateUseCase.m:57-70)  */
#define EEPACC_MAX_CURV     48   /* curvature PWA points    (GenerateUseCase.m:75-92)  */
#define EEPACC_MAX_SLOPE    24   /* road-slope PWA points   (GenerateUseCase.m:97-107) */
#define EEPACC_MAX_STOP     8    /* stop locations          (Settings.m:178)           */
#define EEPACC_MAX_TL       4    /* traffic lights          (Settings.m:188)           */

/* controller kinds (RunOpt_ABMPC.m / RunOpt_FBMPC.m / RunOpt_BLMPC.m / RunOpt_TVMPC.m) */
enum { EEPACC_AB = 0, EEPACC_FB = 1, EEPACC_BL = 2, EEPACC_TV = 3 };

/* iw[1] flags */
#define EEPACC_FLAG_DEVICE_PTR  1   /* arg/res are device pointers: no hidden copies  */

/* return codes */
#define EEPACC_OK               0
#define EEPACC_ERR_ARG          1
#define EEPACC_ERR_NO_DEVICE    2
#define EEPACC_ERR_CUDA         3
#define EEPACC_ERR_MEM          4

/* per-vehicle QP status written to the status output (qpOASES returnValue classes) */
#define EEPACC_QP_OK            0
#define EEPACC_QP_MAXITER       1   /* RET_MAX_NWSR_REACHED                */
#define EEPACC_QP_INFEASIBLE    2   /* RET_*_STOPPED_INFEASIBILITY         */
#define EEPACC_QP_UNBOUNDED     3   /* RET_*_STOPPED_UNBOUNDEDNESS         */
#define EEPACC_QP_NUMERIC       4   /* NaN / factorisation breakdown       */

/* All tunables of Settings.m + SetVehicleParameters.m that reach the hot path. */
typedef struct eepacc_params {
    /* --- vehicle (Functions/Settings/SetVehicleParameters.m, both trees) --- */
    double m, L, h_g, L_f, c_r, R_w, phi;
    double P_m_max, T_m_max, omega_m_r, eta_TF, lambda, mu, g, zeta_a, v_max;
    double F2, p10, p01;            /* Daily coast-down / efficiency-map fit (ABO tree only; 0 for the i3) */
    /* --- objective weights (Settings.m:13-76) --- */
    double W_FB[7];                 /* w_P w_a w_j w_v w_h w_s w_f                          */
    double W_AB[7];                 /* w_FC w_a w_j w_v w_h w_s w_f (w_FC=0 in the MC tree) */
    double W_BL[4];                 /* w_v w_a w_j w_f                                      */
    double W_TV[4];
    double W_NLP[7];                /* w_P w_a w_j w_v w_h w_s w_f (Settings.m:13-28; stage cost L of eepacc_f) */
    /* --- horizon / estimation (Settings.m:96-120) --- */
    double Ts;
    double tConstACC_ego, tConstACC_tar;
    int    N;                       /* must equal EEPACC_N_HOR                              */
    int    n_integrate_plant;       /* RK4 sub-steps of the plant (Settings.m:109)          */
    int    est_ego, est_tar;        /* paramEstSetting, TVestSetting                        */
    int    est_bl, est_tv;          /* BL_trajEstSett, TV_trajEstSett                       */
    int    ab_route_rows;           /* 1: MC CreateQP_AB.m:307-329 rows present; 0: ABO     */
    int    fb_ad_quirk;             /* 1: reproduce RunOpt_FBMPC.m:251-254 (A(k)/D(k) bug)  */
    int    nlp_use_fifth;           /* useFifthOrderFit_NLP (Settings.m:92)                 */
    int    pad0_;
    /* --- use case (Settings.m:141,201-224; CreateQP_FB.m:461-463) --- */
    double s_goal, h_min, tau_min;
    double stopVel, stopRefDist, stopRefVelSlope, TLStopRegionSize, TLstopVel, alpha_TTL;
    double T_hwp, A_hwp, G_hwp;
    double BL_a_low, BL_a_high, BL_j_low, BL_j_high;
    double TV_a_low, TV_a_high, TV_j_low, TV_j_high;
    /* --- power-map fits (Settings.m:227-236) --- */
    double b_quadr[6];
    double b_fifth[21];
    /* --- initial conditions (Settings.m:83-85, 206-207) --- */
    double s_init, v_init, a_minus1_init;
    double cutOffDist;              /* KPI cut-off distance (Settings.m:190, Main.m:136-148) */
} eepacc_params;

/* One scenario's route preview (GenerateUseCase.m output), contiguous so that a CTA
 * pulls it with one bulk copy.  Unused tail entries are ignored (counts below). */
typedef struct eepacc_route {
    double s_speedLim[EEPACC_MAX_SL],   v_speedLim[EEPACC_MAX_SL];
    double s_curv[EEPACC_MAX_CURV],     curvature[EEPACC_MAX_CURV];
    double s_slope[EEPACC_MAX_SLOPE],   slope[EEPACC_MAX_SLOPE];
    double stopLoc[EEPACC_MAX_STOP];
    double TLLoc[EEPACC_MAX_TL][4];     /* [pos, phase, red, green] (Settings.m:187)        */
    double mass_scale;                  /* per-scenario vehicle-mass factor (1 = nominal)   */
    int    n_sl, n_curv, n_slope, n_stop, n_tl, pad_;
} eepacc_route;

/* problem sizes of the dense QP of each controller */
int eepacc_qp_nvar(int ctrl);                              /* 100 / 120 / 40 / 40             */
int eepacc_qp_ncon(int ctrl, const eepacc_params* p);      /* 362|282 / 522 / 262 / 220       */

/* ------------------------------------------------------------------------------------
 * mem objects (casadi_c.h:51-64: checkout/release are NOT thread safe, incref/decref are)
 * ------------------------------------------------------------------------------------ */
/* Bind scenario data to a checked-out mem: copies params + B route blocks to the device,
 * allocates per-vehicle state and resets it to the initial conditions.  `routes` is a host
 * pointer to B blocks.  device < 0 keeps the current CUDA device. */
int eepacc_mem_setup(int mem, int ctrl, const eepacc_params* p, const eepacc_route* routes,
                     long long B, int device);
/* Reset plant / warm-start state to step 0 without re-uploading the scenario. */
int eepacc_mem_reset(int mem);
/* Set the simulation-step counter / clock used by the stand-alone eepacc_qpbuild_*() calls
 * (FB A(k) quirk row, traffic-light phase); the step functions advance both themselves. */
int eepacc_mem_set_step(int mem, int kk, double t0);
/* Upload a whole lead trace: host double[n_steps][B][2] = (s_tv, v_tv) per step.  Enables
 * eepacc_run_*(): the whole closed loop without per-step host involvement. */
int eepacc_mem_set_lead(int mem, const double* lead, long long n_steps);
/* Synthesise the lead traces of a perturbed scenario batch on the device instead of uploading them
 * (replaces the host loop Run_DrivingCycle.m:35-46 + Main.m:88 per scenario).
 * base / hw: host double[n_base] / [n_hw] 2 Hz lead speed samples of the urban / highway cycle (hw may be NULL);
 * pert: host double[B][EEPACC_N_PERT] = scale, delay (steps), initial gap (m), 3 x (amplitude m/s, frequency Hz,
 * phase rad), use-highway flag, 3 spare.  Result: the same device buffer eepacc_mem_set_lead fills. */
int eepacc_mem_synth_lead(int mem, const double* base, long long n_base, const double* hw, long long n_hw,
                          const double* pert, long long n_steps, double tv_length);
/* Route previews built on the device (replaces the host loop GenerateUseCase.m:50-107 + SimplifyPWA.m:14-49 per
 * scenario): what GetUseCase.m / the scenario draw specifies, one block per scenario.  Slopes are given as angles
 * (atan(percent / 100), GenerateUseCase.m:104, evaluated by the caller so that host and device agree bit for bit). */
#define EEPACC_SPEC_MAX_ZONE  6
#define EEPACC_SPEC_MAX_CURVE 12
#define EEPACC_SPEC_MAX_SLOPE 10
typedef struct eepacc_route_spec {
    double zones[EEPACC_SPEC_MAX_ZONE][2];      /* [speed limit km/h, start m]; an empty list = no limit (1e5)      */
    double curves[EEPACC_SPEC_MAX_CURVE][3];    /* [curvature 1/m, s0 m, s1 m]                                      */
    double slopes[EEPACC_SPEC_MAX_SLOPE][3];    /* [angle rad, s0 m, s1 m]                                          */
    double stops[EEPACC_MAX_STOP];
    double tls[EEPACC_MAX_TL][4];               /* [pos, phase, red, green]                                         */
    double mass_scale;
    int    n_zone, n_curve, n_slope, n_stop, n_tl, pad_;
} eepacc_route_spec;
/* Build the B route blocks of a set-up mem from B specs (host pointer) with k5_route_gen_kernel and make them the mem's
 * routes; the closed-loop state is reset.  Fails (EEPACC_ERR_ARG) if a scenario's PWA needs more points than the route
 * block holds; eepacc_last_error() names the first such scenario. */
int eepacc_mem_gen_routes(int mem, const eepacc_route_spec* specs, long long B);
/* Copy the mem's B route blocks back to the host (tests: bit-compare with the host statement). */
int eepacc_mem_fetch_routes(int mem, eepacc_route* out, long long B);
long long eepacc_sizeof_route_spec(void);
/* CUDA stream (cudaStream_t as void*) the mem launches on; default stream if never set. */
int eepacc_mem_set_stream(int mem, void* stream);
/* Number of kernel launches issued through this mem so far (bench.py's gpu_launches). */
int eepacc_mem_device(int mem);                 /* CUDA device the mem is bound to (-1: invalid handle).  Every entry point
                                                  * switches to it and restores the caller's current device on return. */
long long eepacc_mem_launch_count(int mem);
/* Per-vehicle counters accumulated so far: double[B][4] = working-set changes, failed QPs,
 * refactorisations with dropped rows, flips.  Host pointer. */
int eepacc_mem_get_counters(int mem, double* out);
/* Device pointer (double[B][EEPACC_N_KPI]) of the KPI block after eepacc_run(): lets the caller
 * gather KPIs GPU-to-GPU (NCCL) without a host bounce. */
void* eepacc_mem_kpi_device_ptr(int mem);
/* ABI guards for bindings (ctypes / cgo / JNI): sizeof the two PODs as compiled. */
long long eepacc_sizeof_params(void);
long long eepacc_sizeof_route(void);
const char* eepacc_last_error(void);
/* CUDA-event timing of every closed-loop kernel launch on the mem's stream (roofline evidence).
 * tot[4] = summed ms of k2_build, k3_qp, k4_post and the number of steps since the last reset;
 * step_ms[cap] (may be NULL) = per-step batch latency. */
int eepacc_mem_enable_timing(int mem, int on);
int eepacc_mem_get_timing(int mem, double* tot, double* step_ms, long long cap);
/* Developer aid: first call enables a per-iteration solver trace for vehicle 0; later calls fetch
 * out[cap][4] = step length, ratio-test key, |working set|, exchanges of the last solve. */
int eepacc_mem_debug_trace(int mem, double* out, int cap);
/* clock64() cycles of vehicle 0's last solve by solver phase, out[8] (tracing must be enabled). */
int eepacc_mem_debug_phases(int mem, double* out);
/* Closed-loop QP kernel selection.  mode 0 (default): reduced-space active-set kernel (factorisation over the
 * free variables only, capacity fc of them, default 56) with the full-space kernel as fallback for vehicles
 * that do not fit; mode 1: full-space kernel only. */
int eepacc_mem_set_qp_mode(int mem, int mode, int fc);
/* eepacc_run pipelining: the batch advances as `groups` contiguous slices (1..32), each on its own CUDA stream
 * with its own sequence of MPC steps (vehicles never interact; results do not depend on the slicing).
 * 0 = automatic (up to 8 slices of >= 256 vehicles).  With G slices eepacc_mem_get_timing reports one latency
 * per (step, slice), step-major. */
int eepacc_mem_set_groups(int mem, int groups);
/* eepacc_run: 0 (default) = per-step launches over pipelined slices; 1 = after the cold-start step every vehicle
 * runs its remaining MPC steps inside ONE fused kernel launch (no synchronisation between vehicles). */
int eepacc_mem_set_fused(int mem, int on);
int eepacc_mem_get_fused(int mem);
/* Closed-loop QP solver of eepacc_step_* / eepacc_run -- the reference's OPTsettings.solverToUse (MC/Settings.m:112):
 *   EEPACC_SOLVER_QPOASES (default) = solverToUse 1: condensing + dense online active set (K3 / K3r kernels), the path
 *                                     every parity claim of this library is made on;
 *   EEPACC_SOLVER_OCP               = solverToUse 0 / 2 (sparse qpOASES / HPIPM; RunOpt_FBMPC.m:96-121,
 *                                     CreateQP_FB.m:219-308): the same QP kept stage-structured and solved by a
 *                                     Riccati-based primal-dual interior-point method, one vehicle per thread (K6).
 *                                     FBMPC and ABMPC only; cold-started every step; status 0 / 1 (iteration limit) /
 *                                     2 (breakdown); `iter` = interior-point iterations.  It returns the same minimiser
 *                                     wherever that is unique; on a flat optimal face (FBMPC's Fm / Fb split at
 *                                     standstill) it returns the face's analytic centre instead of a vertex.
 * eepacc_mem_set_ocp_options: opts[6] = initial barrier parameter, tolerances on the complementarity, the primal and
 * the stationarity residual, complementarity below which a factorisation breakdown is accepted as converged,
 * iteration limit (NaN keeps the current value; defaults 1e6 (ABMPC) / 1e4 (FBMPC), 1e-11, 1e-9, 1e-10, 1e-9, 80). */
#define EEPACC_SOLVER_QPOASES 0
#define EEPACC_SOLVER_OCP     1
int eepacc_mem_set_solver(int mem, int solver);
int eepacc_mem_get_solver(int mem);
int eepacc_mem_set_ocp_options(int mem, const double* opts);
/* out[7]: [0] duration in ms (CUDA events on the launching stream) of the last fused closed-loop launch, [1] its
 * grid, [2] CTAs resident per SM, [3..6] SM clock cycles summed over all vehicle-steps spent in the build, the QP
 * solve, the hand-over solves and the post stage. */
int eepacc_mem_get_cycle_timing(int mem, double* out);
int eepacc_mem_get_groups(int mem);
/* Developer aid: copy a per-step device buffer of the closed loop to the host (0 sblk, 1 g, 2 lba, 3 uba,
 * 4 x, 5 status, 6 iter, 7 cost, 8 deferred-vehicle count, 9 the lead trace [n_lead][B][2], 10 the working set after the
 * last solve [B][m + n]: -1 lower / +1 upper / 0 inactive, rows then variable bounds); returns the number of doubles written or < 0. */
long long eepacc_mem_debug_fetch(int mem, int what, double* out, long long cap);
/* FP64-FMA peak microbenchmark (register-resident DFMA chains); returns GFLOP/s or <0. */
double eepacc_measure_dfma_peak(int device);

/* ------------------------------------------------------------------------------------
 * CasADi-style function families.  For each F below the library exports
 *   int F(const double** arg, double** res, long long* iw, double* w, int mem);
 *   long long F_n_in(void), F_n_out(void);
 *   const char* F_name_in(long long i), F_name_out(long long i);
 *   const long long* F_sparsity_in(long long i), F_sparsity_out(long long i);   (CCS, dense)
 *   int F_work(long long* sz_arg, long long* sz_res, long long* sz_iw, long long* sz_w);
 *   void F_incref(void), F_decref(void);  int F_checkout(void);  void F_release(int mem);
 *   double F_default_in(long long i);
 * mirroring casadi_c_{n_in,n_out,name_in,name_out,sparsity_in,sparsity_out,work,incref,
 * decref,checkout,release,default_in}_id (casadi_c.h:98-129).
 * ------------------------------------------------------------------------------------ */
#define EEPACC_DECLARE(F)                                                                   \
    int F(const double** arg, double** res, long long* iw, double* w, int mem);             \
    long long F##_n_in(void);                                                               \
    long long F##_n_out(void);                                                              \
    const char* F##_name_in(long long i);                                                   \
    const char* F##_name_out(long long i);                                                  \
    const long long* F##_sparsity_in(long long i);                                          \
    const long long* F##_sparsity_out(long long i);                                         \
    int F##_work(long long* sz_arg, long long* sz_res, long long* sz_iw, long long* sz_w);  \
    void F##_incref(void);                                                                  \
    void F##_decref(void);                                                                  \
    int F##_checkout(void);                                                                 \
    void F##_release(int mem);                                                              \
    double F##_default_in(long long i);

/* K1 -- expression evaluator (RunOpt_NLP.m:225-245,263-288; GetMotorPower_*.m)
 *   eepacc_f      : (x[4]=[s v theta j], u[6]=[Fm Fb xi_v xi_h xi_s xi_f]) -> (xdot[4], L[1])
 *   eepacc_F      : (x0[4], p[6]) -> (xf[4], qf[1])      4 RK4 steps over Ts
 *   eepacc_F_jac  : (x0[4], p[6]) -> (xf[4], qf[1], J[5x10] col-major = d[xf;qf]/d[x0;p])
 *   eepacc_power5 : (Fm[1], rpm[1]) -> (P[1])            quintic surface
 *   eepacc_power2 : (Fm[1], rpm[1]) -> (P[1])            quadratic fit                    */
EEPACC_DECLARE(eepacc_f)
EEPACC_DECLARE(eepacc_F)
EEPACC_DECLARE(eepacc_F_jac)
EEPACC_DECLARE(eepacc_power5)
EEPACC_DECLARE(eepacc_power2)

/* K4 -- plant step (RunPlantModel.m:27-44):  (x[2]=[s v], u[2]=[Fm Fb]) -> (x_next[2])     */
EEPACC_DECLARE(eepacc_plant)

/* K2 -- bounds + QP build + condensing (CreateQP_*.m + TransformToDenseFormulation.m)
 *   in : x0[2]=[s v], a_minus1[1], lead[3]=[s_tv v_tv a_tv], t0[1]
 *   out: h[n x n], g[n], a[m x n], lba[m], uba[m]   (conic scheme names, conic.hpp:170-208) */
EEPACC_DECLARE(eepacc_qpbuild_ab)
EEPACC_DECLARE(eepacc_qpbuild_fb)
EEPACC_DECLARE(eepacc_qpbuild_bl)
EEPACC_DECLARE(eepacc_qpbuild_tv)

/* K3 -- dense warm-started active-set QP, the CasADi conic('qpoases') boundary
 *   (RunOpt_ABMPC.m:252, RunOpt_FBMPC.m:278).  in : h g a lba uba lbx ubx x0 lam_x0 lam_a0
 *   out: x cost lam_a lam_x status iter.  Dimensions follow the mem's controller.        */
EEPACC_DECLARE(eepacc_qpsolve)

/* Fused closed-loop step K5->K4->K2->K3 (+AB force allocation); one call == one iteration
 * of the `for kk = 0:N_sim` loop (RunOpt_FBMPC.m:161-331).
 *   in : lead[2]=[s_tv(k) v_tv(k)]
 *   out: z1[8]=[s v Fm Fb xi_v xi_h xi_s xi_f], a[1], status[1], iter[1]                  */
EEPACC_DECLARE(eepacc_step_ab)
EEPACC_DECLARE(eepacc_step_fb)
EEPACC_DECLARE(eepacc_step_bl)
EEPACC_DECLARE(eepacc_step_tv)

/* Whole closed loop on the device (needs eepacc_mem_set_lead):  in: n_steps[1] (scalar, NOT
 * batched); out: log[n_steps][B][12] = z1[8], a, status, iter, cost (may be NULL);  kpi[B][8]. */
EEPACC_DECLARE(eepacc_run)

/* KPI block per scenario (Main.m:136-148,188-262; RunOpt_FBMPC.m:334-339), what the ranks gather over NCCL:
 * [0] E_opt(end) J   [1] s_opt(end) m   [2] travel time at cutOffDist s   [3] E_opt at cut-off J
 * [4] max a   [5] min a   [6] rms a (over 1..distInd)   [7] failed QPs (sum exitMessage)
 * comfort, Main.m:247-250 (j = diff(a)/Ts over 1..distInd):   [8] max j   [9] min j   [10] rms j
 * constraint violations, counted over the whole run (steps with a slack above 1e-6, the soft constraints of
 * CreateQP_FB.m:211-215):   [11] xi_v (speed incentive)   [12] xi_h (desired headway)   [13] xi_s (safe headway /
 * speed rows)   [14] xi_f (feasibility)
 * headway against the lead trace (finite s_tv only):   [15] steps with gap < h_min   [16] steps with
 * gap < tau_min * v   [17] smallest gap m (+inf without a lead)                                                  */
#define EEPACC_N_KPI 18
#define EEPACC_N_PERT 16

#ifdef __cplusplus
}
#endif
#endif /* EEPACC_H */
