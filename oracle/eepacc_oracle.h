/*
 * eepacc_oracle.h -- CPU ORACLE (test infrastructure, NOT a product path).
 *
 * Plain-C restatement of the reference's per-step MPC hot path (MATLAB sources under
 * /root/reference/{MATLAB_CasADi,ACCMPC-ABO_CasADi}); every function cites the file:line
 * it follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (CUDA) path never does.
 *
 * Parity pins (tests/test_oracle_golden.py):
 *   - plant step, power maps, RK4 map F, QP build + condensing: pinned against the
 *     reference's saved golden .mat data (tests/golden/ npz files).
 *   - QP solve: the reference solver is qpOASES 3.2 inside CasADi 3.6.3's conic plugin
 *     (binary only, absent from /root/reference).  The oracle's own dense parametric
 *     active-set solver (qp_oracle.c) is pinned through the golden first-stage primal
 *     values of the 871 consecutive warm-started QPs of each saved closed-loop run.
 *     Objective values, duals and active sets are NOT stored by the reference:
 *     for those, parity is unpinned by reference data and defined by this oracle.
 */
#ifndef EEPACC_ORACLE_H
#define EEPACC_ORACLE_H

#include "../include/eepacc.h"   /* POD parameter / route structs + enums only */

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_N      EEPACC_N_HOR
#define ORC_NMAX   120          /* max dense variables  (FB: 6*20)           */
#define ORC_MMAX   522          /* max dense rows       (FB: 26*20+2)        */
#define ORC_NZMAX  162          /* max sparse variables (FB: 8*20+2)         */
#define ORC_INF    1.0e20       /* what +-inf bounds are mapped to           */

/* ---------- scalar pieces ---------- */
double orc_interp_pwa(double d, const double* doms, const double* vals, int n);
void   orc_plant(const eepacc_params* p, const eepacc_route* r, const double x[2],
                 const double u[2], double xn[2]);
double orc_power5(const double b[21], double Fm, double rpm);
double orc_power2(const double b[6], double Fm, double rpm);
void   orc_f(const eepacc_params* p, double mass_scale, const double x[4], const double u[6],
             double xdot[4], double* L);
void   orc_F(const eepacc_params* p, double mass_scale, const double x0[4], const double u[6],
             double xf[4], double* qf);
/* J is 5x10 column-major: d[xf;qf]/d[x0;p] */
void   orc_F_jac(const eepacc_params* p, double mass_scale, const double x0[4], const double u[6],
                 double xf[4], double* qf, double J[50]);
void   orc_est_traj(const eepacc_params* p, int estSetting, double s, double v, double a,
                    const double* s_prev_sol, const double* v_prev_sol,
                    double* s_est, double* v_est);
void   orc_bounds(const eepacc_params* p, const eepacc_route* r, int mpctype,
                  const double* s_est, const double* v_est, double t0,
                  double* slope_est, double* v_lim, double* v_stop, double* v_tl, double* v_curv,
                  double* a_min, double* a_max, double* j_min, double* j_max);
void   orc_force_alloc(const eepacc_params* p, const eepacc_route* r, double s_meas,
                       double v_meas, double a_cmd, double* Fm, double* Fb, double* a_real);

/* ---------- QP build + condensing (dense, column-major outputs) ---------- */
typedef struct orc_qp {
    int n, m;                       /* dense sizes */
    double H[ORC_NMAX * ORC_NMAX];  /* col-major n x n */
    double g[ORC_NMAX];
    double A[ORC_MMAX * ORC_NMAX];  /* col-major m x n */
    double lba[ORC_MMAX], uba[ORC_MMAX];
    /* recovery z = Psi x + d, only the rows the loop needs */
    double Psi[ORC_NZMAX * ORC_NMAX]; /* col-major nz x n */
    double d[ORC_NZMAX];
    int nz;
} orc_qp;

/* FB quirk state (RunOpt_FBMPC.m:78-90,247-259) */
typedef struct orc_fb_ad { double A[ORC_N][4]; double D[ORC_N][2]; } orc_fb_ad;
void orc_fb_ad_init(const eepacc_params* p, double mass_scale, orc_fb_ad* ad);

/* builds the dense QP of one step.  `ad` only for FB; k_sim = 1-based simulation step. */
int orc_qpbuild(int ctrl, const eepacc_params* p, const eepacc_route* r,
                double s0, double v0, double a_minus1, double s_tv, double v_tv, double a_tv,
                double t0, const double* s_prev_sol, const double* v_prev_sol,
                orc_fb_ad* ad, int k_sim, orc_qp* qp);

/* ---------- dense parametric active-set QP (qp_oracle.c) ---------- */
typedef struct orc_qpws orc_qpws;
orc_qpws* orc_qpws_new(int n, int m);
void      orc_qpws_free(orc_qpws* w);
void      orc_qpws_reset(orc_qpws* w);          /* forget the warm start */
/* conic scheme (conic.hpp:170-208): min 1/2 x'Hx + g'x, lba<=Ax<=uba, lbx<=x<=ubx.
 * lbx/ubx may be NULL (= +-inf).  Outputs: x[n], cost, lam_a[m], lam_x[n] (CasADi sign
 * convention: >0 at an upper bound).  Returns EEPACC_QP_* status; *iter = working-set changes. */
int orc_qpsolve(orc_qpws* w, const double* H, const double* g, const double* A,
                const double* lba, const double* uba, const double* lbx, const double* ubx,
                double* x, double* cost, double* lam_a, double* lam_x, int* iter);
/* working set after the last solve: ws[m+n] in {-1 lower,0 inactive,+1 upper} */
void orc_qpws_get_ws(const orc_qpws* w, int* ws);

/* ---------- closed loop ---------- */
typedef struct orc_vehicle orc_vehicle;
orc_vehicle* orc_vehicle_new(int ctrl, const eepacc_params* p, const eepacc_route* r);
void         orc_vehicle_free(orc_vehicle* v);
/* one iteration of `for kk=0:N_sim` (RunOpt_*MPC.m); out[12] = z1[8], a, status, iter, cost */
int orc_vehicle_step(orc_vehicle* v, double s_tv_k, double v_tv_k, double* out);
/* last built QP / working set of the vehicle (for parity tests) */
const orc_qp* orc_vehicle_qp(const orc_vehicle* v);
void orc_vehicle_get_ws(const orc_vehicle* v, int* ws);
void orc_vehicle_get_sol(const orc_vehicle* v, double* x, double* lam_a);
/* run n_steps for one vehicle: lead[n_steps][2]; log[n_steps][12]; kpi[EEPACC_N_KPI] */
int orc_run(int ctrl, const eepacc_params* p, const eepacc_route* r, const double* lead,
            int lead_stride, int n_steps, double* log, double* kpi);
void orc_kpi(const eepacc_params* p, double mass_scale, const double* log, int n_steps, double* kpi);
/* B scenarios in a C loop (CPU baseline): routes[B], lead [n_steps][B][2], kpi [B][EEPACC_N_KPI], log_scratch [n_steps][12] */
int orc_run_batch(int ctrl, const eepacc_params* p, const eepacc_route* routes, const double* lead, long long B,
                  int n_steps, double* kpi, double* log_scratch);

int orc_qp_nvar(int ctrl);
int orc_qp_ncon(int ctrl, const eepacc_params* p);

#ifdef __cplusplus
}
#endif
/* nlp_oracle.c: RunOpt_NLP.m's problem functions (objective, constraint rows, bounds) in the reference's order */
double orc_lut_linear(double x, const double* xs, const double* ys, int n);
int    orc_nlp_rows_per_stage(const eepacc_route* r);
int    orc_nlp_eval(const eepacc_params* p, const eepacc_route* r, const double* s_velInc, const double* v_velInc, int n_vi,
                    const double* s_stop, const double* v_stop, int n_stop, const double* s_tv, int N, const double* z,
                    double* f, double* g, double* lbg, double* ubg, double* lbz, double* ubz);

#endif
