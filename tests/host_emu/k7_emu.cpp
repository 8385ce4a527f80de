// k7_emu.cpp -- TEST INFRASTRUCTURE: the NLP solver of csrc/k7_nlp.cuh compiled for the host (g++), one scenario per
// call.  It exists so that the solver can be run and checked on a machine without a GPU (tests/test_nlp_host_emu.py).
// It is never loaded by the product package; the product path (k7_nlp.cu) has no CPU fallback.
#include <cuda_runtime.h>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include "../../eepacc_mpc_casadi_matlab_b200/csrc/k7_nlp.cuh"

extern "C" long long k7_emu_stage_doubles(void) { return K7_STAGE; }

// z_guess may be NULL (built-in start).  z_out: [4 + 10 N], stats: [EEPACC_NLP_N_STATS + 1] (last: objective f)
extern "C" int k7_emu_solve(const eepacc_params* p, const eepacc_route* r, const eepacc_nlp_tables* tb, const double* s_tv,
                            int N, const eepacc_nlp_opts* op, const double* z_guess, double* z_out, double* stats)
{
    K7Ctx c;
    k7_ctx_init(c, p, r, tb, s_tv, N, op->obj_scale);
    K7Work w;
    w.ws = (double*)calloc((size_t)(N + 1) * K7_STAGE, sizeof(double));
    w.stride = 1;
    if (!w.ws) return -1;
    if (z_guess) k7_read_guess(c, w, z_guess); else k7_start(c, w);
    K7Result res;
    memset(&res, 0, sizeof res);
    k7_solve(c, w, *op, res);
    k7_write_z(c, w, z_out);
    stats[0] = res.status; stats[1] = res.iters; stats[2] = res.kkt; stats[3] = res.infeas; stats[4] = res.compl_;
    stats[5] = res.mu; stats[6] = res.delta; stats[7] = res.backtracks; stats[8] = res.f;
    free(w.ws);
    return 0;
}

// one stage evaluated in isolation (values and the Jacobian rows), for the derivative tests
extern "C" int k7_emu_stage(const eepacc_params* p, const eepacc_route* r, const eepacc_nlp_tables* tb, const double* s_tv,
                            int N, int k, const double* x, const double* u, double* xn, double* cost, double* rows,
                            double* jac /* [(4 + 1 + nrow)][10] or NULL */)
{
    K7Ctx c;
    k7_ctx_init(c, p, r, tb, s_tv, N, 1.0);
    if (!jac) {
        K7Stage<double> o;
        k7_stage<double, true>(c, k, x, u, o);
        for (int i = 0; i < 4; ++i) xn[i] = o.xn[i];
        *cost = o.cost;
        for (int i = 0; i < c.nrow; ++i) rows[i] = o.row[i];
        return c.nrow;
    }
    K7Dual X[4], U[6];
    k7_seed(x, u, X, U);
    K7Stage<K7Dual> o;
    k7_stage<K7Dual, true>(c, k, X, U, o);
    for (int i = 0; i < 4; ++i) { xn[i] = o.xn[i].v; memcpy(jac + 10 * i, o.xn[i].d, 80); }
    *cost = o.cost.v; memcpy(jac + 40, o.cost.d, 80);
    for (int i = 0; i < c.nrow; ++i) { rows[i] = o.row[i].v; memcpy(jac + 10 * (5 + i), o.row[i].d, 80); }
    return c.nrow;
}
