/* eepacc_nlp.h -- C-ABI of the batched NLP benchmark: the full-route optimal-control problem of
 * MATLAB_CasADi/RunOpt_NLP.m:204-510 (multiple shooting over the whole driving cycle, solved there by
 * CasADi nlpsol('ipopt'), 620 s for one scenario in savedNLPsol.mat:tSolve), for B independent scenarios at once.
 *
 * What the entry point replaces:   RunOpt_NLP.m:499-510
 *     prob   = struct('f', J, 'x', z, 'g', g);  solver = nlpsol('solver', 'ipopt', prob, opts);
 *     sol    = solver('x0', z0, 'lbx', lbz, 'ubx', ubz, 'lbg', lbg, 'ubg', ubg);   z_opt = full(sol.x);
 * The calling convention is the CasADi-generated one of casadi_fun.c (arg / res / iw / w / mem) with a leading batch
 * dimension, as in eepacc.h; input / output names follow CasADi's nlpsol scheme (x0, p -> x, f).  The bounds
 * lbx / ubx / lbg / ubg are not arguments: RunOpt_NLP.m builds them from the settings (:190-199, :340-497), and so does
 * the kernel, from eepacc_params + eepacc_route + eepacc_nlp_tables.
 *
 * z layout (the reference's, RunOpt_NLP.m:308-372 and :533-543), per scenario, n_z = 4 + 10 * n_steps doubles:
 *     [ X_0 | U_0 X_1 | U_1 X_2 | ... | U_{N-1} X_N ],   X = (s, v, theta, j),   U = (Fm, Fb, xi_v, xi_h, xi_s, xi_f).
 *
 * There is no CPU fallback: without a CUDA device eepacc_nlp_setup fails with EEPACC_ERR_NO_DEVICE.
 */
#ifndef EEPACC_NLP_H
#define EEPACC_NLP_H
#include "eepacc.h"

#ifdef __cplusplus
extern "C" {
#endif

#define EEPACC_NLP_MAX_VI    48   /* points of the velocity-incentive PWA (RunOpt_NLP.m:160-183)              */
#define EEPACC_NLP_MAX_STOPV 24   /* points of the stop velocity profile, 3 per stop (RunOpt_NLP.m:86-112)    */
#define EEPACC_NLP_N_STATS   8    /* res[2] columns: status, iterations, kkt error, primal infeasibility,
                                     complementarity, final mu, regularisation used, backtracking steps        */

/* per-scenario status (stats column 0).  ACCEPTABLE: the iteration stalled at the round-off floor of the merit function at
 * the final barrier parameter -- ipopt's "Solved To Acceptable Level" (RunOpt_NLP.m:253-254 sets those options). */
enum { EEPACC_NLP_OK = 0, EEPACC_NLP_MAXITER = 1, EEPACC_NLP_LINESEARCH = 2, EEPACC_NLP_NUMERIC = 3, EEPACC_NLP_ACCEPTABLE = 4 };

/* Set-up products of RunOpt_NLP.m:86-184 that are tables, built once per scenario on the host
 * (eepacc_mpc_casadi_matlab_b200/nlp.py: nlp_tables) and read by the kernel through linear lookups with CasADi's
 * `interpolant('LUT','linear')` semantics (linear extrapolation from the end cells). */
typedef struct eepacc_nlp_tables {
    double s_velInc[EEPACC_NLP_MAX_VI], v_velInc[EEPACC_NLP_MAX_VI];       /* :160-183 velIncentiveLookup      */
    double s_stop[EEPACC_NLP_MAX_STOPV], v_stop[EEPACC_NLP_MAX_STOPV];     /* :86-112  stopMaxVelLookup        */
    int    n_vi, n_stop;
} eepacc_nlp_tables;

typedef struct eepacc_nlp_opts {
    double tol;          /* scaled KKT tolerance (ipopt tol, default 1e-8)                                   */
    double mu0;          /* initial barrier parameter                                                        */
    double obj_scale;    /* objective scaling applied inside the solver (f is reported unscaled)             */
    double mu_min;       /* last barrier parameter (default 1e-9 = tol / 10).  At small mu the multipliers of tight
                            headway rows carry the round-off of positions of ~3e3 m, lambda^2 ulp(s) / mu ~ 0.1 at
                            1e-9: the reported KKT error has that floor, the primal solution does not (k7_nlp.cuh)  */
    double acceptable_kkt; /* a stalled iterate ends the solve as ACCEPTABLE below this scaled KKT error (default 1) */
    int    max_iter;     /* RunOpt_NLP.m:250 opts.ipopt.max_iter = NLPmaxIter                                */
    int    pad_;
} eepacc_nlp_opts;

int  eepacc_nlp_checkout(void);                    /* casadi_c_checkout_id: a memory slot, or a negative id  */
void eepacc_nlp_release(int mem);
/* Bind B scenarios of n_steps intervals (t_sim / Ts) to a device.  p: one parameter set; routes, tables: [B]. */
int  eepacc_nlp_setup(int mem, const eepacc_params* p, const eepacc_route* routes, const eepacc_nlp_tables* tables,
                      long long B, int n_steps, int device);
int  eepacc_nlp_set_opts(int mem, const eepacc_nlp_opts* o);
void eepacc_nlp_default_opts(eepacc_nlp_opts* o);
/* arg[0] = x0 [B][n_z] initial guess in the z layout above, or NULL: the built-in start (a car-following roll-out);
 * arg[1] = p  [B][n_steps + 1]: lead-vehicle rear-bumper position s_tv (OPTsettings.s_tv, RunOpt_NLP.m:46)
 * res[0] = x  [B][n_z];  res[1] = f [B] (objective J);  res[2] = stats [B][EEPACC_NLP_N_STATS]
 * iw[0] = B (must equal the set-up batch), iw[1] = EEPACC_FLAG_DEVICE_PTR when arg / res are device pointers.
 * Returns 0, or an EEPACC_ERR_* code; per-scenario solver status is in stats, never in the return code. */
int  eepacc_nlp(const double** arg, double** res, long long* iw, double* w, int mem);
long long   eepacc_nlp_n_in(void);
long long   eepacc_nlp_n_out(void);
const char* eepacc_nlp_name_in(long long i);
const char* eepacc_nlp_name_out(long long i);
int  eepacc_nlp_work(long long* sz_arg, long long* sz_res, long long* sz_iw, long long* sz_w);
long long   eepacc_nlp_nz(int n_steps);            /* 4 + 10 * n_steps                                       */
long long   eepacc_nlp_workspace_bytes(long long B, int n_steps);
double      eepacc_nlp_last_ms(int mem);           /* device time of the last eepacc_nlp call (CUDA events)  */
const char* eepacc_nlp_last_error(void);
long long   eepacc_nlp_sizeof_tables(void);

#ifdef __cplusplus
}
#endif
#endif
