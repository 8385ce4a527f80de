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
#include <algorithm>
using std::min;
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
static inline double __shfl_xor_sync(unsigned, double v, int) { return v; }
template <class T> static inline T __ldg(const T* p) { return *p; }
#include "../../eepacc_mpc_casadi_matlab_b200/csrc/k3r_qp.cu"
#include "../../eepacc_mpc_casadi_matlab_b200/csrc/k2_build.cuh"
#include "../../eepacc_mpc_casadi_matlab_b200/csrc/k4_post.cuh"
#include "../../eepacc_mpc_casadi_matlab_b200/csrc/k6_ocp.cuh"
#include <vector>
#include <cstdio>

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

// ---- closed loop of ONE vehicle through the product's device code (K2 build -> K3r reduced-space QP, a solve
// that outgrows `fc` repeated at capacity n -> K4 post), single-threaded on the host.  Mirrors
// eepacc_capi.cu:device_step_slice.  log: [n_steps][12]; ws (optional): [n_steps][m + n] working set after each
// solve (-1 lower / 0 / +1 upper), the layout of orc_vehicle_get_ws.
template <int NU, int NB>
static void emu_qp(const eep_qp_io& io, unsigned char* buf) { k3r_qp_kernel<NU, NB>(io, buf); }

extern "C" int k3r_emu_run(int ctrl, const eepacc_params* p, const eepacc_route* route, const double* lead, int n_steps,
                           int fc, double* log, signed char* ws_out, int* stats)
{
    const int NH = EEP_N;
    const int nu = (ctrl == EEPACC_FB) ? 6 : (ctrl == EEPACC_AB) ? 5 : 2;
    const int n = NH * nu;
    int m;
    {   // rows of the controller: same count as the library (eepacc_qp_ncon): probe the row generator
        m = (ctrl == EEPACC_FB) ? 26 * NH + 2 : (ctrl == EEPACC_AB) ? ((p->ab_route_rows ? 18 : 14) * NH + 2) : (ctrl == EEPACC_BL) ? 13 * NH + 2 : 11 * NH;
    }
    const int mt = n + m;
    if (fc > n) fc = n;
    std::vector<double> s_prev(1), v_prev(1), Fm_prev(1), Fb_prev(1), a_m1(1), v_tv_meas(1), fbA(NH), fbD(NH), prev_sol(2 * (NH + 1)),
        x0(2), g(n), lba(m), uba(m), sblk(EEP_SBLK), dyn(2 * NH), wx(n), wy(n), x(n), cost(1), status(1), iter(1), kpi_acc(ACC_W),
        hscr(4 * (size_t)mt + 1), consts(96);
    std::vector<short> widx(n); std::vector<signed char> wside(n); std::vector<int> wnA(1), warm(1), dlist(2), dcnt(1);
    std::vector<unsigned long long> tab(m + 1), counters(4);
    const int rc = k3r_build_rowtab(ctrl, *p, m, tab.data(), consts.data());
    if (rc) return 100 + rc;
    std::vector<unsigned char> buf(eep_qpr_smem_bytes(n, m, fc) + 64), bufn(eep_qpr_smem_bytes(n, m, n) + 64);
    BuildSmem* S = new BuildSmem;
    double t0 = 0.0;
    for (int kk = 0; kk < n_steps; ++kk) {
        eep_build_io b; memset(&b, 0, sizeof b);
        b.B = 1; b.ctrl = ctrl; b.m = m; b.mode = 0; b.kk = kk; b.t0 = t0; b.routes = route; b.route_stride = 1;
        b.s_prev = s_prev.data(); b.v_prev = v_prev.data(); b.Fm_prev = Fm_prev.data(); b.Fb_prev = Fb_prev.data();
        b.a_minus1 = a_m1.data(); b.v_tv_meas = v_tv_meas.data(); b.fbA = fbA.data(); b.fbD = fbD.data();
        b.prev_sol = prev_sol.data(); b.lead = lead ? lead + 2 * (size_t)kk : nullptr;
        b.x0 = x0.data(); b.g = g.data(); b.lba = lba.data(); b.uba = uba.data(); b.sblk = sblk.data(); b.dyn = dyn.data();
        k2_vehicle<1>(*p, b, 0, *S);
        eep_qp_io q; memset(&q, 0, sizeof q);
        q.B = 1; q.n = n; q.m = m; q.ctrl = ctrl; q.g = g.data(); q.lba = lba.data(); q.uba = uba.data(); q.sblk = sblk.data();
        q.wx = wx.data(); q.wy = wy.data(); q.widx = widx.data(); q.wside = wside.data(); q.wnA = wnA.data(); q.warm = warm.data();
        q.x = x.data(); q.cost = cost.data(); q.status = status.data(); q.iter = iter.data(); q.counters = counters.data();
        q.fc = fc; q.defer_list = dlist.data(); q.defer_cnt = dcnt.data(); q.rowtab = tab.data(); q.rowc = consts.data(); q.hscr = hscr.data();
        dcnt[0] = 0;
        if (ctrl == EEPACC_FB) emu_qp<6, 2>(q, buf.data()); else if (ctrl == EEPACC_AB) emu_qp<5, 1>(q, buf.data()); else emu_qp<2, 1>(q, buf.data());
        if (dcnt[0]) {
            if (stats) { stats[0] += 1; if (kk == 0) stats[1] += 1; }
            q.fc = n; dcnt[0] = 0;
            if (ctrl == EEPACC_FB) emu_qp<6, 2>(q, bufn.data()); else if (ctrl == EEPACC_AB) emu_qp<5, 1>(q, bufn.data()); else emu_qp<2, 1>(q, bufn.data());
            if (dcnt[0]) { delete S; return -2; }
        }
        if (ws_out) {
            signed char* w = ws_out + (size_t)kk * mt;
            memset(w, 0, mt);
            for (int i = 0; i < n; ++i) if (widx[i] >= 0) w[widx[i]] = wside[i];
        }
        eep_post_io po; memset(&po, 0, sizeof po);
        po.B = 1; po.ctrl = ctrl; po.n = n; po.kk = kk; po.routes = route; po.route_stride = 1;
        po.x = x.data(); po.x0 = x0.data(); po.dyn = dyn.data(); po.status = status.data(); po.iter = iter.data(); po.cost = cost.data();
        po.s_prev = s_prev.data(); po.v_prev = v_prev.data(); po.Fm_prev = Fm_prev.data(); po.Fb_prev = Fb_prev.data();
        po.prev_sol = prev_sol.data(); po.log = log + (size_t)kk * EEPACC_LOG_W; po.kpi_acc = kpi_acc.data();
        k4_vehicle(*p, po, 0);
        t0 += p->Ts;
    }
    delete S;
    return 0;
}

// ---- unit entry points of the product's device functions (compiled for the host): the estimator in all its modes
// and the per-stage route / comfort bounds, for CPU tests against the oracle ----
extern "C" void k3r_emu_est_traj(const eepacc_params* p, int est_setting, double s, double v, double a, const double* sp,
                                 const double* vp, double* s_est, double* v_est)
{
    eep_est_traj(*p, est_setting, s, v, a, sp, vp, s_est, v_est);
}
// out[N][9] = slope, v_lim, v_stop, v_tl, v_curv, a_min, a_max, j_min, j_max per stage
extern "C" void k3r_emu_bounds(const eepacc_params* p, const eepacc_route* r, int mpctype, const double* s_est, const double* v_est,
                               double t0, double* out)
{
    for (int i = 0; i < EEP_N; ++i) {
        eep_stage_bounds B;
        eep_bounds_stage(*p, r, mpctype, i, s_est[i], v_est[i], t0, B);
        double* o = out + 9 * i;
        o[0] = B.slope; o[1] = B.v_lim; o[2] = B.v_stop; o[3] = B.v_tl; o[4] = B.v_curv; o[5] = B.a_min; o[6] = B.a_max; o[7] = B.j_min; o[8] = B.j_max;
    }
}



// debug: objective of the structured formulation at a given control vector u (forward simulation of the deviations)
template <int NU, int NB>
static double k6_eval_cost(const k6::Args& A, const double* u, double* viol)
{
    const int NH = EEP_N; constexpr int NX = k6::Dim<NU, NB>::NX, NZ = k6::Dim<NU, NB>::NZ;
    double ds[EEP_N + 1], dv[EEP_N + 1]; ds[0] = dv[0] = 0;
    for (int k = 0; k < NH; ++k) { double drv = 0; for (int c = 0; c < NB; ++c) drv += u[k * NU + c];
        ds[k + 1] = ds[k] + A.T * dv[k] + A.M0 * drv; dv[k + 1] = A.alpha[k] * dv[k] + A.M1 * drv; }
    double cost = 0, vmax = 0; const k6::SbView sb{ A.sb, 1 };
    const int rps = A.m / NH, nterm = A.m - rps * NH;
    for (int k = 0; k <= NH; ++k) {
        double zk[NZ]; zk[0] = ds[k]; zk[1] = dv[k]; zk[2] = k > 0 ? dv[k - 1] : 0;
        for (int c = 0; c < NB; ++c) zk[3 + c] = k > 0 ? u[(k - 1) * NU + c] : 0;
        for (int c = 0; c < NU; ++c) zk[NX + c] = k < NH ? u[k * NU + c] : 0;
        const int c0 = (k < NH) ? k * rps : NH * rps, c1 = (k < NH) ? c0 + rps : c0 + nterm;
        for (int c = c0; c < c1; ++c) { k6::SparseRow r; k6::load_row<NU, NB>(A.rowtab[c], A.rowc, sb, k, r);
            double az = 0; for (int e = 0; e < r.nnz; ++e) az += r.val[e] * zk[r.idx[e]];
            if (A.lba[c] - az > vmax) vmax = A.lba[c] - az; if (az - A.uba[c] > vmax) vmax = az - A.uba[c]; }
        cost += A.qz[k] * dv[k];
        if (k == NH) break;
        double core[1 + NB], e[1 + NB]; core[0] = dv[k]; e[0] = zk[2];
        for (int c = 0; c < NB; ++c) { core[1 + c] = u[k * NU + c]; e[1 + c] = zk[3 + c]; }
        const double* Dc = A.sb + EEP_SB_DC + k * 9; const double* Cc = A.sb + EEP_SB_CC + k * 9;
        for (int a = 0; a <= NB; ++a) for (int b = 0; b <= NB; ++b) cost += 0.5 * core[a] * Dc[a * 3 + b] * core[b] + core[a] * Cc[a * 3 + b] * e[b];
        for (int c = 0; c < NU; ++c) { const double uu = u[k * NU + c]; if (c >= NB) cost += 0.5 * A.sb[EEP_SB_DG + k * 4 + (c - NB)] * uu * uu; cost += A.qz[NH + 1 + k * NU + c] * uu; }
    }
    *viol = vmax; return cost;
}

// ---- closed loop of ONE vehicle with the stage-structured interior-point solver K6 (k6_ocp.cuh) in the solver's
// place: K2 build -> K6 -> K4 post.  opts = { mu0, eps_mu, eps_p, eps_d, maxit, warm, check }.  With check != 0
// every QP is also solved by the reduced-space active-set kernel (warm-started on the same sequence of QPs) and
// diag[n_steps][6] = { max |u - u_dense| over the first stage, over all stages, SIGNED relative objective difference,
// K6 iterations, K6 status, dense status } is filled.
extern "C" int k6_emu_run(int ctrl, const eepacc_params* p, const eepacc_route* route, const double* lead, int n_steps,
                          const double* opts, double* log, double* diag)
{
    const int NH = EEP_N;
    if (ctrl != EEPACC_FB && ctrl != EEPACC_AB) return -10;
    const int nu = (ctrl == EEPACC_FB) ? 6 : 5, nb = (ctrl == EEPACC_FB) ? 2 : 1;
    const int n = NH * nu;
    const int m = (ctrl == EEPACC_FB) ? 26 * NH + 2 : ((p->ab_route_rows ? 18 : 14) * NH + 2);
    const int mt = n + m;
    const int check = opts[6] != 0.0;
    std::vector<double> s_prev(1), v_prev(1), Fm_prev(1), Fb_prev(1), a_m1(1), v_tv_meas(1), fbA(NH), fbD(NH), prev_sol(2 * (NH + 1)),
        x0(2), g(n), lba(m), uba(m), sblk(EEP_SBLK), dyn(2 * NH), wx(n), wy(n), x(n), xd(n), cost(1), costd(1), status(1), statusd(1),
        iter(1), iterd(1), kpi_acc(ACC_W), hscr(4 * (size_t)mt + 1), consts(96), qz(NH + 1 + n),
        work((size_t)k6::work_items(n, m, nu, nb));
    std::vector<short> widx(n); std::vector<signed char> wside(n); std::vector<int> wnA(1), warm(1), dlist(2), dcnt(1);
    std::vector<unsigned long long> tab(m + 1), counters(4);
    const int rc = k3r_build_rowtab(ctrl, *p, m, tab.data(), consts.data());
    if (rc) return 100 + rc;
    std::vector<unsigned char> bufn(eep_qpr_smem_bytes(n, m, n) + 64);
    BuildSmem* S = new BuildSmem;
    double t0 = 0.0;
    int have = 0;
    for (int kk = 0; kk < n_steps; ++kk) {
        eep_build_io b; memset(&b, 0, sizeof b);
        b.B = 1; b.ctrl = ctrl; b.m = m; b.mode = 0; b.kk = kk; b.t0 = t0; b.routes = route; b.route_stride = 1;
        b.s_prev = s_prev.data(); b.v_prev = v_prev.data(); b.Fm_prev = Fm_prev.data(); b.Fb_prev = Fb_prev.data();
        b.a_minus1 = a_m1.data(); b.v_tv_meas = v_tv_meas.data(); b.fbA = fbA.data(); b.fbD = fbD.data();
        b.prev_sol = prev_sol.data(); b.lead = lead ? lead + 2 * (size_t)kk : nullptr;
        b.x0 = x0.data(); b.g = g.data(); b.lba = lba.data(); b.uba = uba.data(); b.sblk = sblk.data(); b.dyn = dyn.data();
        b.qz = qz.data();
        k2_vehicle<1>(*p, b, 0, *S);
        k6::Args A;
        A.sb = sblk.data(); A.alpha = dyn.data(); A.qz = qz.data(); A.lba = lba.data(); A.uba = uba.data();
        A.rowtab = tab.data(); A.rowc = consts.data(); A.work = work.data(); A.ws = 1; A.m = m;
        A.T = p->Ts;
        A.M0 = (ctrl == EEPACC_FB) ? 0.0 : 0.5 * p->Ts * p->Ts;
        A.M1 = (ctrl == EEPACC_FB) ? p->Ts / p->lambda / sblk[EEP_SB_MASS] : p->Ts;
        A.mu0 = opts[0]; A.eps_mu = opts[1]; A.eps_p = opts[2]; A.eps_d = opts[3]; A.maxit = (int)opts[4]; A.eps_accept = opts[7]; A.gamma = opts[8]; A.tau0 = opts[9];
        A.warm = (opts[5] != 0.0 && have) ? 1 : 0;
        int it = 0; double cst = 0.0;
        const int st = (ctrl == EEPACC_FB) ? k6::solve<6, 2>(A, x.data(), cst, it) : k6::solve<5, 1>(A, x.data(), cst, it);
        have = 1;
        cost[0] = cst; status[0] = st; iter[0] = it;
        if (check) {
            eep_qp_io q; memset(&q, 0, sizeof q);
            q.B = 1; q.n = n; q.m = m; q.ctrl = ctrl; q.g = g.data(); q.lba = lba.data(); q.uba = uba.data(); q.sblk = sblk.data();
            q.wx = wx.data(); q.wy = wy.data(); q.widx = widx.data(); q.wside = wside.data(); q.wnA = wnA.data(); q.warm = warm.data();
            q.x = xd.data(); q.cost = costd.data(); q.status = statusd.data(); q.iter = iterd.data(); q.counters = counters.data();
            q.fc = n; q.defer_list = dlist.data(); q.defer_cnt = dcnt.data(); q.rowtab = tab.data(); q.rowc = consts.data(); q.hscr = hscr.data();
            dcnt[0] = 0;
            if (ctrl == EEPACC_FB) emu_qp<6, 2>(q, bufn.data()); else emu_qp<5, 1>(q, bufn.data());
            double e1 = 0, ea = 0;
            for (int j = 0; j < n; ++j) { const double e = fabs(x[j] - xd[j]); if (j < nu && e > e1) e1 = e; if (e > ea) ea = e; }
            if (getenv("K6_DEBUG")) { double v1, v2; double c1 = (ctrl == EEPACC_FB) ? k6_eval_cost<6, 2>(A, xd.data(), &v1) : k6_eval_cost<5, 1>(A, xd.data(), &v1);
                double c2 = (ctrl == EEPACC_FB) ? k6_eval_cost<6, 2>(A, x.data(), &v2) : k6_eval_cost<5, 1>(A, x.data(), &v2);
                printf("step %d: dense cost %.12e  structured-formula cost at dense x %.12e (viol %.2e) | k6 cost %.12e at k6 x %.12e (viol %.2e)\n", kk, costd[0], c1, v1, cst, c2, v2); }
            double* dg = diag + (size_t)kk * 6;
            dg[0] = e1; dg[1] = ea; dg[2] = (cst - costd[0]) / (1.0 + fabs(costd[0]));     /* signed: < 0 = K6 found the lower objective */ dg[3] = it; dg[4] = st; dg[5] = statusd[0];
        } else if (diag) { double* dg = diag + (size_t)kk * 6; dg[0] = dg[1] = dg[2] = 0; dg[3] = it; dg[4] = st; dg[5] = 0; }
        eep_post_io po; memset(&po, 0, sizeof po);
        po.B = 1; po.ctrl = ctrl; po.n = n; po.kk = kk; po.routes = route; po.route_stride = 1;
        po.x = x.data(); po.x0 = x0.data(); po.dyn = dyn.data(); po.status = status.data(); po.iter = iter.data(); po.cost = cost.data();
        po.s_prev = s_prev.data(); po.v_prev = v_prev.data(); po.Fm_prev = Fm_prev.data(); po.Fb_prev = Fb_prev.data();
        po.prev_sol = prev_sol.data(); po.log = log + (size_t)kk * EEPACC_LOG_W; po.kpi_acc = kpi_acc.data();
        k4_vehicle(*p, po, 0);
        t0 += p->Ts;
    }
    delete S;
    return 0;
}
