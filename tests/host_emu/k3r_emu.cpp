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
#include <vector>

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
