// eepacc_capi.cu -- the C-ABI of include/eepacc.h: CasADi-codegen style entry points
//   int F(const double** arg, double** res, long long* iw, double* w, int mem)
// (casadi_c_eval_id, casadi_c.h:128-129; called from ACCMPC-ABO_CasADi/casadi_fun.c:170) with a
// leading batch dimension, plus the mem objects that own device-resident per-vehicle state.
// No CPU fallback: every F() launches sm_100a kernels or fails.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <mutex>
#include <string>
#include <vector>

#include "eepacc_kernels.h"

#define NH EEPACC_N_HOR
#define ACC_W EEP_ACC_W
#define MAX_MEM 64

namespace {

thread_local std::string g_err;
int fail(int code, const char* what, cudaError_t e = cudaSuccess)
{
    char buf[512];
    if (e != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    else snprintf(buf, sizeof buf, "%s", what);
    g_err = buf;
    return code;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(EEPACC_ERR_CUDA, #call, e_); } while (0)

#define MAX_GROUPS 32

struct DBuf {
    void* p = nullptr; size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) return fail(EEPACC_ERR_MEM, "cudaMalloc", e);
        cap = bytes;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct Mem {
    bool used = false, ready = false;
    int ctrl = 0, n = 0, m = 0, device = 0;
    long long B = 0;
    cudaStream_t stream = nullptr;
    eepacc_params p;
    long long launches = 0;
    int kk = 0;
    double t0 = 0.0;
    long long n_lead = 0;
    DBuf routes, s_prev, v_prev, Fm_prev, Fb_prev, a_minus1, v_tv_meas, fbA, fbD, prev_sol, x0, dyn;
    DBuf g, lba, uba, sblk;     // closed loop is matrix-free: no dense H / A is ever stored
    DBuf wx, wy, widx, wside, wnA, warm, xsol, cost, status, iter, counters;
    DBuf kpi_acc, kpi, lead, lead_step, logrow, dbg, synth_in;
    DBuf rowtab, rowc;          // packed constraint rows of the controller for the reduced-space QP kernel
    DBuf defer_list, defer_cnt; // vehicles the reduced-space QP kernel hands to the full-space kernel
    DBuf hscr;                  // [B][4 (n + m)] homotopy state of the reduced-space QP kernel
    DBuf defer_list2;           // second hand-over list: reduced-space kernel at full capacity -> full-space kernel
    int qp_fc2 = 0;             // capacity of the second-level reduced-space launch (0: none)
    int qp_fbgrid = 16;         // CTAs of a hand-over launch past the cold-start step
    int fused = 0;              // eepacc_run: 1 = steps >= 1 in the fused closed-loop kernel (k5_cycle_kernel)
    DBuf cyc_counter, cyc_big;  // its vehicle hand-out counter and the work area of its hand-over solves
    cudaEvent_t cyc_ev[2] = { nullptr, nullptr };   // around the last fused launch (always recorded: two events per cycle)
    int cyc_grid = 0, cyc_per_sm = 0;
    int qp_mode = 0;            // 0: reduced-space kernel + full-space fallback, 1: full-space kernel only
    int solver = EEPACC_SOLVER_QPOASES;   // closed-loop QP solver: dense active set (K3/K3r) or stage-structured interior point (K6)
    DBuf qz, ocp_work;          // K6 only: un-condensed gradient [B][N+1+n], per-vehicle work space [items][B]
    double ocp_opts[6] = { 1e6, 1e-11, 1e-9, 1e-10, 1e-9, 80 };   // mu0, eps_mu, eps_p, eps_d, eps_accept, maxit
    int qp_fc = 56;             // capacity (free variables) of the reduced factorisation (measured: 48 -> 56 +1 % AB, +3 % FB; 64, 72 slower)
    long long deferred_total = 0;
    DBuf in[12], out[8];        // staging for host-pointer calls
    bool timing = false;        // CUDA-event timing of every kernel launch (bench.py roofline leg)
    std::vector<cudaEvent_t> evs;
    int ev_steps = 0, ev_groups = 1;
    // eepacc_run pipelines the batch as n_groups contiguous slices, each with its own stream and its own
    // sequence of MPC steps: vehicles never interact, so a slice whose slowest vehicle needs many
    // working-set changes no longer idles the SMs the other slices can use
    int n_groups = 1;
    cudaStream_t gstream[MAX_GROUPS] = {};
    cudaEvent_t gfork = nullptr, gjoin[MAX_GROUPS] = {};
    void release_all()
    {
        DBuf* all[] = { &routes, &s_prev, &v_prev, &Fm_prev, &Fb_prev, &a_minus1, &v_tv_meas, &fbA, &fbD, &prev_sol,
                        &x0, &dyn, &g, &lba, &uba, &sblk, &wx, &wy, &widx, &wside, &wnA, &warm, &xsol, &cost,
                        &status, &iter, &counters, &kpi_acc, &kpi, &lead, &lead_step, &logrow, &dbg, &defer_list, &defer_cnt, &rowtab, &rowc, &hscr, &defer_list2, &cyc_counter, &cyc_big, &synth_in, &qz, &ocp_work };
        for (DBuf* b : all) b->release();
        for (auto& b : in) b.release();
        for (auto& b : out) b.release();
        for (auto e : evs) cudaEventDestroy(e);
        evs.clear(); ev_steps = 0;
        for (int g = 0; g < MAX_GROUPS; ++g) {
            if (gstream[g]) cudaStreamDestroy(gstream[g]);
            if (gjoin[g]) cudaEventDestroy(gjoin[g]);
            gstream[g] = nullptr; gjoin[g] = nullptr;
        }
        if (gfork) cudaEventDestroy(gfork);
        gfork = nullptr;
        for (int e = 0; e < 2; ++e) { if (cyc_ev[e]) cudaEventDestroy(cyc_ev[e]); cyc_ev[e] = nullptr; }
        n_lead = 0; kk = 0; t0 = 0.0; launches = 0; deferred_total = 0; cyc_grid = 0; cyc_per_sm = 0;
        ready = false;
    }
};

Mem g_mem[MAX_MEM];
std::mutex g_mu;
int g_refcount = 0;
int g_last_ctrl = EEPACC_FB;
int g_last_m = 522;

Mem* get_mem(int mem)
{
    if (mem < 0 || mem >= MAX_MEM || !g_mem[mem].used) { fail(EEPACC_ERR_ARG, "invalid mem handle"); return nullptr; }
    return &g_mem[mem];
}

int nvar(int ctrl) { return ctrl == EEPACC_FB ? 120 : ctrl == EEPACC_AB ? 100 : 40; }
int ncon(int ctrl, const eepacc_params* p)
{
    switch (ctrl) {
    case EEPACC_FB: return 522;
    case EEPACC_AB: return (p && !p->ab_route_rows) ? 282 : 362;
    case EEPACC_BL: return 262;
    default: return 220;
    }
}

int reset_state(Mem* M)
{
    const long long B = M->B;
    CK(cudaMemsetAsync(M->s_prev.p, 0, B * 8, M->stream));
    CK(cudaMemsetAsync(M->v_prev.p, 0, B * 8, M->stream));
    CK(cudaMemsetAsync(M->Fm_prev.p, 0, B * 8, M->stream));
    CK(cudaMemsetAsync(M->Fb_prev.p, 0, B * 8, M->stream));
    CK(cudaMemsetAsync(M->v_tv_meas.p, 0, B * 8, M->stream));
    CK(cudaMemsetAsync(M->prev_sol.p, 0, B * 2 * (NH + 1) * 8, M->stream));
    CK(cudaMemsetAsync(M->warm.p, 0, B * sizeof(int), M->stream));
    CK(cudaMemsetAsync(M->wnA.p, 0, B * sizeof(int), M->stream));
    CK(cudaMemsetAsync(M->counters.p, 0, B * 4 * sizeof(unsigned long long), M->stream));
    CK(cudaMemsetAsync(M->kpi_acc.p, 0, B * ACC_W * 8, M->stream));
    // a_minus1 and the FB A/D rows start from their Settings.m / RunOpt_FBMPC.m:78-90 values
    std::vector<double> am1((size_t)B, M->p.a_minus1_init);
    CK(cudaMemcpyAsync(M->a_minus1.p, am1.data(), B * 8, cudaMemcpyHostToDevice, M->stream));
    CK(cudaStreamSynchronize(M->stream));
    M->kk = 0; M->t0 = 0.0;
    return EEPACC_OK;
}

// copy-in helper: returns device pointer for input `src` (host or device)
int stage_in(Mem* M, int slot, const double* src, size_t count, bool dev, const double** out)
{
    if (!src) { *out = nullptr; return 0; }
    if (dev) { *out = src; return 0; }
    if (M->in[slot].ensure(count * 8)) return EEPACC_ERR_MEM;
    CK(cudaMemcpyAsync(M->in[slot].p, src, count * 8, cudaMemcpyHostToDevice, M->stream));
    *out = M->in[slot].as<double>();
    return 0;
}
int stage_out(Mem* M, int slot, double* dst, size_t count, bool dev, double** out)
{
    if (!dst) { *out = nullptr; return 0; }
    if (dev) { *out = dst; return 0; }
    if (M->out[slot].ensure(count * 8)) return EEPACC_ERR_MEM;
    *out = M->out[slot].as<double>();
    return 0;
}
int copy_out(Mem* M, int slot, double* dst, size_t count, bool dev)
{
    if (!dst || dev) return 0;
    CK(cudaMemcpyAsync(dst, M->out[slot].p, count * 8, cudaMemcpyDeviceToHost, M->stream));
    return 0;
}

// The entry points switch to the mem's device and put the caller's current device back when they return
// (also on the error paths: the guard's destructor does it).
struct DeviceGuard {
    int prev = -1;
    int enter(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev == dev) { prev = -1; return 0; }          // already there: nothing to switch, nothing to restore
        return (int)cudaSetDevice(dev);
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
struct CallCtx { Mem* M; long long B; bool dev; DeviceGuard guard; };
int begin_call(int mem, long long* iw, CallCtx& c)
{
    Mem* M = get_mem(mem);
    if (!M) return EEPACC_ERR_ARG;
    if (!M->ready) return fail(EEPACC_ERR_ARG, "mem not set up (call eepacc_mem_setup)");
    if (!iw) return fail(EEPACC_ERR_ARG, "iw is NULL (iw[0] = batch, iw[1] = flags)");
    CK((cudaError_t)c.guard.enter(M->device));
    c.M = M; c.B = iw[0]; c.dev = (iw[1] & EEPACC_FLAG_DEVICE_PTR) != 0;
    if (c.B < 0) return fail(EEPACC_ERR_ARG, "negative batch");
    return 0;
}
int end_call(CallCtx& c)
{
    CK(cudaGetLastError());
    if (!c.dev) CK(cudaStreamSynchronize(c.M->stream));
    return EEPACC_OK;
}

eep_build_io make_build_io(Mem* M, int mode)
{
    eep_build_io io;
    memset(&io, 0, sizeof io);
    io.B = M->B; io.ctrl = M->ctrl; io.m = M->m; io.mode = mode; io.kk = M->kk; io.t0 = M->t0;
    io.routes = M->routes.as<eepacc_route>(); io.route_stride = 1;
    io.s_prev = M->s_prev.as<double>(); io.v_prev = M->v_prev.as<double>();
    io.Fm_prev = M->Fm_prev.as<double>(); io.Fb_prev = M->Fb_prev.as<double>();
    io.a_minus1 = M->a_minus1.as<double>(); io.v_tv_meas = M->v_tv_meas.as<double>();
    io.fbA = M->fbA.as<double>(); io.fbD = M->fbD.as<double>();
    io.prev_sol = M->prev_sol.as<double>();
    io.x0 = M->x0.as<double>();
    io.H = nullptr; io.g = M->g.as<double>(); io.A = nullptr;
    io.lba = M->lba.as<double>(); io.uba = M->uba.as<double>();
    io.sblk = M->sblk.as<double>();
    io.dyn = M->dyn.as<double>();
    return io;
}
eep_qp_io make_qp_io(Mem* M)
{
    eep_qp_io io;
    memset(&io, 0, sizeof io);
    io.B = M->B; io.n = M->n; io.m = M->m; io.ctrl = M->ctrl;
    io.H = nullptr; io.g = M->g.as<double>(); io.A = nullptr;
    io.lba = M->lba.as<double>(); io.uba = M->uba.as<double>();
    io.sblk = M->sblk.as<double>();
    io.wx = M->wx.as<double>(); io.wy = M->wy.as<double>();
    io.widx = M->widx.as<short>(); io.wside = M->wside.as<signed char>();
    io.wnA = M->wnA.as<int>(); io.warm = M->warm.as<int>();
    io.x = M->xsol.as<double>(); io.cost = M->cost.as<double>();
    io.status = M->status.as<double>(); io.iter = M->iter.as<double>();
    io.counters = M->counters.as<unsigned long long>();
    io.dbg = M->dbg.as<double>(); io.dbg_cap = M->dbg.p ? 4096 : 0;
    return io;
}

// Restrict the kernel argument blocks to the vehicles [v0, v0 + nb): every per-vehicle array is
// indexed from its base pointer, so a slice is the same block with shifted pointers.
template <class T> static inline T* shift(T* p, size_t k) { return p ? p + k : p; }
static void slice_io(eep_build_io& io, size_t v0, long long nb, int m, int n)
{
    io.B = nb;
    io.routes = shift(io.routes, v0 * (size_t)io.route_stride);
    io.s_prev = shift(io.s_prev, v0); io.v_prev = shift(io.v_prev, v0);
    io.Fm_prev = shift(io.Fm_prev, v0); io.Fb_prev = shift(io.Fb_prev, v0);
    io.a_minus1 = shift(io.a_minus1, v0); io.v_tv_meas = shift(io.v_tv_meas, v0);
    io.fbA = shift(io.fbA, v0 * NH); io.fbD = shift(io.fbD, v0 * NH);
    io.prev_sol = shift(io.prev_sol, v0 * 2 * (NH + 1));
    io.lead = shift(io.lead, v0 * 2);
    io.x0 = shift(io.x0, v0 * 2);
    io.g = shift(io.g, v0 * n); io.lba = shift(io.lba, v0 * m); io.uba = shift(io.uba, v0 * m);
    io.sblk = shift(io.sblk, v0 * EEP_SBLK);
    io.dyn = shift(io.dyn, v0 * 2 * NH);
}
static void slice_io(eep_qp_io& io, size_t v0, long long nb)
{
    const size_t n = io.n, m = io.m;
    io.B = nb;
    io.g = shift(io.g, v0 * n); io.lba = shift(io.lba, v0 * m); io.uba = shift(io.uba, v0 * m);
    io.sblk = shift(io.sblk, v0 * EEP_SBLK);
    io.wx = shift(io.wx, v0 * n); io.wy = shift(io.wy, v0 * n);
    io.widx = shift(io.widx, v0 * n); io.wside = shift(io.wside, v0 * n);
    io.wnA = shift(io.wnA, v0); io.warm = shift(io.warm, v0);
    io.x = shift(io.x, v0 * n); io.cost = shift(io.cost, v0);
    io.status = shift(io.status, v0); io.iter = shift(io.iter, v0);
    io.counters = shift(io.counters, v0 * 4);
    if (v0) { io.dbg = nullptr; io.dbg_cap = 0; }       // the trace follows vehicle 0 of the batch
}
static void slice_io(eep_post_io& io, size_t v0, long long nb)
{
    const size_t n = io.n;
    io.B = nb;
    io.routes = shift(io.routes, v0 * (size_t)io.route_stride);
    io.x = shift(io.x, v0 * n); io.x0 = shift(io.x0, v0 * 2); io.dyn = shift(io.dyn, v0 * 2 * NH);
    io.status = shift(io.status, v0); io.iter = shift(io.iter, v0); io.cost = shift(io.cost, v0);
    io.s_prev = shift(io.s_prev, v0); io.v_prev = shift(io.v_prev, v0);
    io.Fm_prev = shift(io.Fm_prev, v0); io.Fb_prev = shift(io.Fb_prev, v0);
    io.prev_sol = shift(io.prev_sol, v0 * 2 * (NH + 1));
    io.log = shift(io.log, v0 * EEPACC_LOG_W);
    io.kpi_acc = shift(io.kpi_acc, v0 * ACC_W);
    io.lead = shift(io.lead, v0 * 2);
}

static eep_post_io make_post_io(Mem* M, int kk, double* log_dev)
{
    eep_post_io pio;
    memset(&pio, 0, sizeof pio);
    pio.B = M->B; pio.ctrl = M->ctrl; pio.n = M->n; pio.kk = kk;
    pio.routes = M->routes.as<eepacc_route>(); pio.route_stride = 1;
    pio.x = M->xsol.as<double>(); pio.x0 = M->x0.as<double>(); pio.dyn = M->dyn.as<double>();
    pio.status = M->status.as<double>(); pio.iter = M->iter.as<double>(); pio.cost = M->cost.as<double>();
    pio.s_prev = M->s_prev.as<double>(); pio.v_prev = M->v_prev.as<double>();
    pio.Fm_prev = M->Fm_prev.as<double>(); pio.Fb_prev = M->Fb_prev.as<double>();
    pio.prev_sol = M->prev_sol.as<double>();
    pio.log = log_dev;
    pio.kpi_acc = M->kpi_acc.as<double>();
    return pio;
}

// One closed-loop step, K5/K4/K2 -> K3 -> post, of the vehicles [v0, v0 + nb) on stream st (slice g of
// the batch) at MPC step kk / time t0.  lead_dev: [B][2], log_dev: [B][EEPACC_LOG_W] of the whole batch.
// ev: four events (build start, build end, QP end, post end) or null.
int device_step_slice(Mem* M, int g, size_t v0, long long nb, cudaStream_t st, int kk, double t0,
                      const double* lead_dev, double* log_dev, cudaEvent_t* ev)
{
    eep_build_io bio = make_build_io(M, 0);
    bio.kk = kk; bio.t0 = t0;
    bio.lead = lead_dev;
    slice_io(bio, v0, nb, M->m, M->n);
    if (M->solver == EEPACC_SOLVER_OCP) bio.qz = M->qz.as<double>() + v0 * (size_t)(NH + 1 + M->n);
    if (ev) CK(cudaEventRecord(ev[0], st));
    int rc = eep_launch_build(&M->p, &bio, st);
    if (ev) CK(cudaEventRecord(ev[1], st));
    if (rc) return fail(EEPACC_ERR_CUDA, "k2_build launch", (cudaError_t)rc);
    eep_qp_io qio = make_qp_io(M);
    slice_io(qio, v0, nb);
    bool need_full = true;      // does anything remain for the full-space kernel?
    if (M->solver == EEPACC_SOLVER_OCP) {
        // stage-structured interior-point solver, one vehicle per thread (k6_ocp.cu); cold-started every step, no solver state
        eep_ocp_io oio;
        memset(&oio, 0, sizeof oio);
        oio.B = nb; oio.n = M->n; oio.m = M->m; oio.ctrl = M->ctrl;
        oio.sblk = qio.sblk; oio.lba = qio.lba; oio.uba = qio.uba;
        oio.dyn = M->dyn.as<double>() + v0 * 2 * (size_t)NH;
        oio.qz = M->qz.as<double>() + v0 * (size_t)(NH + 1 + M->n);
        oio.rowtab = M->rowtab.as<unsigned long long>(); oio.rowc = M->rowc.as<double>();
        oio.work = M->ocp_work.as<double>() + v0; oio.work_stride = M->B;
        oio.x = qio.x; oio.cost = qio.cost; oio.status = qio.status; oio.iter = qio.iter; oio.counters = qio.counters;
        oio.mu0 = M->ocp_opts[0]; oio.eps_mu = M->ocp_opts[1]; oio.eps_p = M->ocp_opts[2]; oio.eps_d = M->ocp_opts[3];
        oio.eps_accept = M->ocp_opts[4]; oio.maxit = (int)M->ocp_opts[5];
        rc = eep_launch_ocp(&M->p, &oio, st);
        if (rc) return fail(EEPACC_ERR_CUDA, "k6_ocp launch", (cudaError_t)rc);
        M->launches += 1;
        need_full = false;
    } else if (M->qp_mode == 0) {
        // Hand-over launches (second level, full space) take a whole SM's shared memory per CTA: a CTA of theirs
        // cannot start before an SM has drained, and the slice waits for the whole grid.  Past the cold-start
        // step the hand-over lists hold well under 1 % of the vehicles, so a handful of CTAs walks them.
        qio.todo_grid = (kk == 0) ? 148 : M->qp_fbgrid;
        // reduced-space kernel first; whatever it cannot hold goes to the full-space kernel
        int* cnt = M->defer_cnt.as<int>() + 4 * g;
        CK(cudaMemsetAsync(cnt, 0, 2 * sizeof(int), st));
        qio.fc = M->qp_fc;
        qio.defer_list = M->defer_list.as<int>() + v0; qio.defer_cnt = cnt;      // indices relative to v0
        qio.rowtab = M->rowtab.as<unsigned long long>(); qio.rowc = M->rowc.as<double>();
        qio.hscr = M->hscr.as<double>() + v0 * 4 * (size_t)(M->n + M->m);
        rc = eep_launch_qpr(&M->p, &qio, st);
        if (rc) return fail(EEPACC_ERR_CUDA, "k3r_qp launch", (cudaError_t)rc);
        qio.todo_list = qio.defer_list; qio.todo_cnt = cnt;
        M->launches += 1;
        if (M->qp_fc2 > M->qp_fc) {
            // second level: the same kernel with (almost) every variable allowed to be free, one CTA per SM,
            // over the hand-over list (the cold-start step and the rare wide working sets); what even that
            // cannot hold goes on to the full-space kernel
            eep_qp_io q2 = qio;
            q2.fc = M->qp_fc2;
            q2.defer_list = M->defer_list2.as<int>() + v0; q2.defer_cnt = cnt + 1;
            rc = eep_launch_qpr(&M->p, &q2, st);
            if (rc) return fail(EEPACC_ERR_CUDA, "k3r_qp second-level launch", (cudaError_t)rc);
            qio.todo_list = q2.defer_list; qio.todo_cnt = cnt + 1;
            M->launches += 1;
            if (M->qp_fc2 >= M->n) need_full = false;     // every variable may be free: the second level never hands over
        }
        if (M->qp_fc >= M->n) need_full = false;
    }
    if (need_full) {
        rc = eep_launch_qp(&M->p, &qio, st);
        if (rc) return fail(EEPACC_ERR_CUDA, "k3_qp launch", (cudaError_t)rc);
        M->launches += 1;
    }
    if (ev) CK(cudaEventRecord(ev[2], st));
    eep_post_io pio = make_post_io(M, kk, log_dev);
    pio.lead = (M->ctrl == EEPACC_TV) ? nullptr : lead_dev;      // headway KPIs (TVMPC has no target vehicle)
    slice_io(pio, v0, nb);
    rc = eep_launch_post(&M->p, &pio, st);
    if (rc) return fail(EEPACC_ERR_CUDA, "k4_post launch", (cudaError_t)rc);
    if (ev) CK(cudaEventRecord(ev[3], st));
    M->launches += 2;
    return EEPACC_OK;
}

static int timing_events(Mem* M, size_t need)
{
    while (M->evs.size() < need) { cudaEvent_t e; CK(cudaEventCreate(&e)); M->evs.push_back(e); }
    return 0;
}

// the whole batch, one step, on the mem's stream (eepacc_step_* and the single-slice eepacc_run)
int device_step(Mem* M, const double* lead_dev, double* log_dev)
{
    cudaEvent_t* ev = nullptr;
    if (M->timing) {
        if (timing_events(M, (size_t)(M->kk + 1) * 4)) return EEPACC_ERR_CUDA;
        ev = &M->evs[(size_t)M->kk * 4];
        M->ev_steps = M->kk + 1; M->ev_groups = 1;
    }
    int rc = device_step_slice(M, 0, 0, M->B, M->stream, M->kk, M->t0, lead_dev, log_dev, ev);
    if (rc) return rc;
    M->kk += 1;
    M->t0 += M->p.Ts;
    return EEPACC_OK;
}

// ---------------- metadata ----------------
struct Port { const char* name; int rows, cols; };   // rows < 0: resolved from the last set-up controller
enum { DIM_N = -1, DIM_M = -2, DIM_NN = -3, DIM_MN = -4 };

const long long* dense_ccs(int rows, int cols)
{
    static std::mutex mu;
    static std::vector<std::vector<long long>> cache;
    std::lock_guard<std::mutex> lk(mu);
    for (auto& v : cache) if (v[0] == rows && v[1] == cols) return v.data();
    std::vector<long long> v;
    v.push_back(rows); v.push_back(cols);
    for (int c = 0; c <= cols; ++c) v.push_back((long long)c * rows);
    for (int c = 0; c < cols; ++c) for (int r = 0; r < rows; ++r) v.push_back(r);
    cache.push_back(std::move(v));
    return cache.back().data();
}
void resolve(const Port& p, int ctrl, int m, int& rows, int& cols)
{
    const int n = nvar(ctrl);
    rows = p.rows; cols = p.cols;
    if (p.rows == DIM_N) { rows = n; cols = 1; }
    else if (p.rows == DIM_M) { rows = m; cols = 1; }
    else if (p.rows == DIM_NN) { rows = n; cols = n; }
    else if (p.rows == DIM_MN) { rows = m; cols = n; }
}

}  // namespace

extern "C" {

const char* eepacc_last_error(void) { return g_err.c_str(); }
int eepacc_qp_nvar(int ctrl) { return nvar(ctrl); }
int eepacc_qp_ncon(int ctrl, const eepacc_params* p) { return ncon(ctrl, p); }
long long eepacc_sizeof_params(void) { return (long long)sizeof(eepacc_params); }
long long eepacc_sizeof_route(void) { return (long long)sizeof(eepacc_route); }

static int checkout_impl(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    for (int i = 0; i < MAX_MEM; ++i)
        if (!g_mem[i].used) { g_mem[i] = Mem(); g_mem[i].used = true; return i; }
    fail(EEPACC_ERR_MEM, "no free mem slot");
    return -1;
}
static void release_impl(int mem)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (mem < 0 || mem >= MAX_MEM || !g_mem[mem].used) return;
    DeviceGuard dg_; if (g_mem[mem].ready) dg_.enter(g_mem[mem].device);
    g_mem[mem].release_all();
    g_mem[mem].used = false;
}

int eepacc_mem_setup(int mem, int ctrl, const eepacc_params* p, const eepacc_route* routes, long long B, int device)
{
    Mem* M = get_mem(mem);
    if (!M) return EEPACC_ERR_ARG;
    if (!p || B <= 0) return fail(EEPACC_ERR_ARG, "eepacc_mem_setup: bad arguments");      // routes == NULL: built later by eepacc_mem_gen_routes
    if (ctrl < EEPACC_AB || ctrl > EEPACC_TV) return fail(EEPACC_ERR_ARG, "unknown controller");
    if (p->N != NH) return fail(EEPACC_ERR_ARG, "params.N must equal EEPACC_N_HOR (20)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail(EEPACC_ERR_NO_DEVICE, "no CUDA device (there is no CPU fallback)");
    if (device < 0) { CK(cudaGetDevice(&device)); }
    if (device >= ndev) return fail(EEPACC_ERR_NO_DEVICE, "device index out of range");
    DeviceGuard dg_; CK((cudaError_t)dg_.enter(device));
    M->release_all();
    M->device = device; M->ctrl = ctrl; M->p = *p; M->B = B;
    M->n = nvar(ctrl); M->m = ncon(ctrl, p);
    g_last_ctrl = ctrl; g_last_m = M->m;
    const size_t n = M->n, m = M->m;
    struct { DBuf* b; size_t bytes; } plan[] = {
        { &M->routes, (size_t)B * sizeof(eepacc_route) },
        { &M->s_prev, (size_t)B * 8 }, { &M->v_prev, (size_t)B * 8 }, { &M->Fm_prev, (size_t)B * 8 },
        { &M->Fb_prev, (size_t)B * 8 }, { &M->a_minus1, (size_t)B * 8 }, { &M->v_tv_meas, (size_t)B * 8 },
        { &M->fbA, (size_t)B * NH * 8 }, { &M->fbD, (size_t)B * NH * 8 },
        { &M->prev_sol, (size_t)B * 2 * (NH + 1) * 8 }, { &M->x0, (size_t)B * 2 * 8 }, { &M->dyn, (size_t)B * 2 * NH * 8 },
        { &M->g, (size_t)B * n * 8 }, { &M->sblk, (size_t)B * EEP_SBLK * 8 },
        { &M->lba, (size_t)B * m * 8 }, { &M->uba, (size_t)B * m * 8 },
        { &M->wx, (size_t)B * n * 8 }, { &M->wy, (size_t)B * n * 8 }, { &M->widx, (size_t)B * n * 2 },
        { &M->wside, (size_t)B * n }, { &M->wnA, (size_t)B * 4 }, { &M->warm, (size_t)B * 4 },
        { &M->xsol, (size_t)B * n * 8 }, { &M->cost, (size_t)B * 8 }, { &M->status, (size_t)B * 8 },
        { &M->iter, (size_t)B * 8 }, { &M->counters, (size_t)B * 4 * 8 },
        { &M->kpi_acc, (size_t)B * ACC_W * 8 }, { &M->kpi, (size_t)B * EEPACC_N_KPI * 8 },
        { &M->lead_step, (size_t)B * 2 * 8 }, { &M->logrow, (size_t)B * EEPACC_LOG_W * 8 },
        { &M->defer_list, (size_t)(B + 1) * 4 }, { &M->defer_cnt, 16 * MAX_GROUPS },
        { &M->rowtab, (size_t)(m + 1) * 8 }, { &M->rowc, 96 * 8 },
        { &M->hscr, (size_t)B * 4 * (size_t)(n + m) * 8 }, { &M->defer_list2, (size_t)(B + 1) * 4 },
    };
    for (auto& e : plan) if (e.b->ensure(e.bytes)) return EEPACC_ERR_MEM;
    if (routes) CK(cudaMemcpyAsync(M->routes.p, routes, (size_t)B * sizeof(eepacc_route), cudaMemcpyHostToDevice, M->stream));
    else CK(cudaMemsetAsync(M->routes.p, 0, (size_t)B * sizeof(eepacc_route), M->stream));      // to be built by eepacc_mem_gen_routes
    // FB dynamics rows before the first step (RunOpt_FBMPC.m:78-90)
    {
        std::vector<double> fa((size_t)B * NH), fd((size_t)B * NH);
        for (long long b = 0; b < B; ++b) {
            double mass = p->m * (routes ? routes[b].mass_scale : 1.0);      // (K2 rewrites these rows at closed-loop step 0 from the device routes)
            for (int k = 0; k < NH; ++k) {
                fa[b * NH + k] = 1 - 2 * p->Ts * p->zeta_a * p->v_init * p->v_init / p->lambda / mass;
                fd[b * NH + k] = p->Ts / p->lambda / mass * (p->zeta_a * p->v_init * p->v_init);
            }
        }
        CK(cudaMemcpyAsync(M->fbA.p, fa.data(), fa.size() * 8, cudaMemcpyHostToDevice, M->stream));
        CK(cudaMemcpyAsync(M->fbD.p, fd.data(), fd.size() * 8, cudaMemcpyHostToDevice, M->stream));
        CK(cudaStreamSynchronize(M->stream));
    }
    {
        int rc = eep_qpr_build_rowtab(ctrl, p, M->m, M->rowtab.as<unsigned long long>(), M->rowc.as<double>(), M->stream);
        if (rc) return fail(EEPACC_ERR_CUDA, "constraint-row table of the reduced-space QP kernel", rc < 100000 ? (cudaError_t)rc : cudaSuccess);
    }
    // developer knobs (A/B runs): EEPACC_QP_MODE=1 forces the full-space kernel, EEPACC_QP_FC sets the capacity
    M->qp_mode = 0; M->qp_fc = 56;
    if (const char* e = getenv("EEPACC_QP_MODE")) M->qp_mode = atoi(e) ? 1 : 0;
    if (const char* e = getenv("EEPACC_QP_FC")) { int v = atoi(e); if (v >= 8 && v <= EEPACC_MAX_NVAR) M->qp_fc = v; }
    if (M->qp_fc > M->n) M->qp_fc = M->n;
    {
        // largest capacity whose shared memory still fits one CTA per SM (227 KB opt-in limit)
        int f2 = M->n;
        while (f2 > M->qp_fc && eep_qpr_smem_bytes(M->n, M->m, f2) > 227 * 1024) --f2;
        // OFF by default: the cold-start step of FBMPC sits on a flat direction of an indefinite objective, where the
        // path through the working sets decides which KKT point is reached; only the full-space kernel walks the
        // reference path there (golden run of the ABO tree: Fm = 6131 N at step 0).  It also measured no faster.
        M->qp_fc2 = 0;
        if (const char* e = getenv("EEPACC_QP_FC2")) { int v = atoi(e); M->qp_fc2 = (v > M->qp_fc && v <= f2) ? v : 0; }
        if (const char* e = getenv("EEPACC_QP_FBGRID")) { int v = atoi(e); if (v >= 1 && v <= 148) M->qp_fbgrid = v; }
        M->solver = EEPACC_SOLVER_QPOASES;
        M->fused = 0;           // measured slower than the sliced per-step path (profiles/README.md): opt-in
        if (const char* e = getenv("EEPACC_FUSED")) M->fused = atoi(e) ? 1 : 0;
    }
    M->ready = true;
    {
        int G = 0;
        if (const char* e = getenv("EEPACC_GROUPS")) G = atoi(e);
        if (G < 0 || G > MAX_GROUPS) G = 0;
        if (eepacc_mem_set_groups(mem, G)) return EEPACC_ERR_ARG;
    }
    return reset_state(M);
}

/* Closed-loop QP kernel selection: mode 0 = reduced-space kernel (capacity fc free variables, 8..n) with
 * the full-space kernel as fallback for the vehicles that do not fit; mode 1 = full-space kernel only. */
int eepacc_mem_set_qp_mode(int mem, int mode, int fc)
{
    Mem* M = get_mem(mem);
    if (!M || !M->ready) return fail(EEPACC_ERR_ARG, "mem not set up");
    if (mode < 0 || mode > 1) return fail(EEPACC_ERR_ARG, "mode must be 0 or 1");
    if (fc <= 0) fc = 56;
    if (fc < 8) fc = 8;
    if (fc > M->n) fc = M->n;
    M->qp_mode = mode; M->qp_fc = fc;
    if (M->qp_fc2 <= fc) M->qp_fc2 = 0;
    return EEPACC_OK;
}

int eepacc_mem_reset(int mem)
{
    Mem* M = get_mem(mem);
    if (!M || !M->ready) return fail(EEPACC_ERR_ARG, "mem not set up");
    DeviceGuard dg_; CK((cudaError_t)dg_.enter(M->device));
    return reset_state(M);
}

int eepacc_mem_set_step(int mem, int kk, double t0)
{
    Mem* M = get_mem(mem);
    if (!M) return EEPACC_ERR_ARG;
    M->kk = kk; M->t0 = t0;
    return EEPACC_OK;
}

int eepacc_mem_set_lead(int mem, const double* lead, long long n_steps)
{
    Mem* M = get_mem(mem);
    if (!M || !M->ready) return fail(EEPACC_ERR_ARG, "mem not set up");
    if (!lead || n_steps <= 0) return fail(EEPACC_ERR_ARG, "bad lead trace");
    DeviceGuard dg_; CK((cudaError_t)dg_.enter(M->device));
    size_t bytes = (size_t)n_steps * M->B * 2 * 8;
    if (M->lead.ensure(bytes)) return EEPACC_ERR_MEM;
    CK(cudaMemcpyAsync(M->lead.p, lead, bytes, cudaMemcpyHostToDevice, M->stream));
    CK(cudaStreamSynchronize(M->stream));
    M->n_lead = n_steps;
    return EEPACC_OK;
}

int eepacc_mem_synth_lead(int mem, const double* base, long long n_base, const double* hw, long long n_hw,
                          const double* pert, long long n_steps, double tv_length)
{
    Mem* M = get_mem(mem);
    if (!M || !M->ready) return fail(EEPACC_ERR_ARG, "mem not set up");
    if (!base || n_base <= 0 || !pert || n_steps < 2 || (hw && n_hw <= 0))
        return fail(EEPACC_ERR_ARG, "bad lead synthesis arguments");
    DeviceGuard dg_; CK((cudaError_t)dg_.enter(M->device));
    const size_t B = (size_t)M->B;
    if (M->lead.ensure((size_t)n_steps * B * 2 * 8)) return EEPACC_ERR_MEM;
    const size_t nb = (size_t)n_base * 8, nh = hw ? (size_t)n_hw * 8 : 0, np_ = B * EEPACC_N_PERT * 8;
    if (M->synth_in.ensure(nb + nh + np_)) return EEPACC_ERR_MEM;
    char* d = (char*)M->synth_in.p;
    CK(cudaMemcpyAsync(d, base, nb, cudaMemcpyHostToDevice, M->stream));
    if (hw) CK(cudaMemcpyAsync(d + nb, hw, nh, cudaMemcpyHostToDevice, M->stream));
    CK(cudaMemcpyAsync(d + nb + nh, pert, np_, cudaMemcpyHostToDevice, M->stream));
    if (eep_launch_lead_synth((long long)B, (int)n_steps, M->p.Ts, tv_length, (const double*)d, (int)n_base,
                              hw ? (const double*)(d + nb) : nullptr, (int)n_hw, (const double*)(d + nb + nh),
                              M->lead.as<double>(), M->stream))
        return fail(EEPACC_ERR_CUDA, "k5_lead_synth_kernel launch");
    M->launches += 1;
    CK(cudaStreamSynchronize(M->stream));
    M->n_lead = n_steps;
    return EEPACC_OK;
}

int eepacc_mem_gen_routes(int mem, const eepacc_route_spec* specs, long long B)
{
    Mem* M = get_mem(mem);
    if (!M || !M->ready) return fail(EEPACC_ERR_ARG, "mem not set up");
    if (!specs || B != M->B) return fail(EEPACC_ERR_ARG, "eepacc_mem_gen_routes: B must equal the batch the mem was set up for");
    DeviceGuard dg_; CK((cudaError_t)dg_.enter(M->device));
    const size_t bytes = (size_t)B * sizeof(eepacc_route_spec);
    if (M->synth_in.ensure(bytes + 16)) return EEPACC_ERR_MEM;
    char* d = (char*)M->synth_in.p;
    int* err = (int*)(d + bytes);
    const int big = 0x7fffffff;
    CK(cudaMemcpyAsync(d, specs, bytes, cudaMemcpyHostToDevice, M->stream));
    CK(cudaMemcpyAsync(err, &big, sizeof(int), cudaMemcpyHostToDevice, M->stream));
    if (eep_launch_route_gen(B, (const eepacc_route_spec*)d, M->routes.as<eepacc_route>(), err, M->stream))
        return fail(EEPACC_ERR_CUDA, "k5_route_gen_kernel launch");
    M->launches += 1;
    int e = 0;
    CK(cudaMemcpyAsync(&e, err, sizeof(int), cudaMemcpyDeviceToHost, M->stream));
    CK(cudaStreamSynchronize(M->stream));
    if (e != big) {
        static char msg[160];
        const char* what[] = { "", "speed-limit", "curvature", "slope", "stop / traffic-light" };
        snprintf(msg, sizeof msg, "eepacc_mem_gen_routes: scenario %d needs more %s points than the route block holds", e / 8, what[e % 8 <= 4 ? e % 8 : 0]);
        return fail(EEPACC_ERR_ARG, msg);
    }
    return reset_state(M);
}

int eepacc_mem_fetch_routes(int mem, eepacc_route* out, long long B)
{
    Mem* M = get_mem(mem);
    if (!M || !M->ready || !out || B != M->B) return fail(EEPACC_ERR_ARG, "eepacc_mem_fetch_routes: bad arguments");
    DeviceGuard dg_; CK((cudaError_t)dg_.enter(M->device));
    CK(cudaStreamSynchronize(M->stream));
    CK(cudaMemcpy(out, M->routes.p, (size_t)B * sizeof(eepacc_route), cudaMemcpyDeviceToHost));
    return EEPACC_OK;
}

long long eepacc_sizeof_route_spec(void) { return (long long)sizeof(eepacc_route_spec); }

int eepacc_mem_set_stream(int mem, void* stream)
{
    Mem* M = get_mem(mem);
    if (!M) return EEPACC_ERR_ARG;
    M->stream = (cudaStream_t)stream;
    return EEPACC_OK;
}

int eepacc_mem_device(int mem)
{
    Mem* M = get_mem(mem);
    return M ? M->device : -1;
}

long long eepacc_mem_launch_count(int mem)
{
    Mem* M = get_mem(mem);
    return M ? M->launches : -1;
}

int eepacc_mem_get_counters(int mem, double* out)
{
    Mem* M = get_mem(mem);
    if (!M || !M->ready || !out) return fail(EEPACC_ERR_ARG, "bad arguments");
    DeviceGuard dg_; CK((cudaError_t)dg_.enter(M->device));
    std::vector<unsigned long long> h((size_t)M->B * 4);
    CK(cudaMemcpyAsync(h.data(), M->counters.p, h.size() * 8, cudaMemcpyDeviceToHost, M->stream));
    CK(cudaStreamSynchronize(M->stream));
    for (size_t i = 0; i < h.size(); ++i) out[i] = (double)h[i];
    return EEPACC_OK;
}

/* device pointer of the per-vehicle KPI block [B][EEPACC_N_KPI] after eepacc_run (for NCCL gathers) */
void* eepacc_mem_kpi_device_ptr(int mem)
{
    Mem* M = get_mem(mem);
    return (M && M->ready) ? M->kpi.p : nullptr;
}

double eepacc_measure_dfma_peak(int device) { return eep_dfma_peak(device); }

/* Developer aid: clock64() cycles of vehicle 0's last solve by phase (needs eepacc_mem_debug_trace enabled):
 * 0 load/norms, 1 warm QR, 2 refactor R, 3 aux-QP setup, 4 KKT step, 5 A*dx + ratio tests + step,
 * 6 working-set change, 7 refinement + outputs; out[8..15] KKT-step sub-phases, out[16..19] ratio-test sub-phases. out[20]. */
int eepacc_mem_debug_phases(int mem, double* out)
{
    Mem* M = get_mem(mem);
    if (!M || !M->ready || !M->dbg.p || !out) return fail(EEPACC_ERR_ARG, "tracing not enabled");
    DeviceGuard dg_; CK((cudaError_t)dg_.enter(M->device));
    CK(cudaMemcpy(out, (double*)M->dbg.p + 4096 * 4, 28 * 8, cudaMemcpyDeviceToHost));
    return EEPACC_OK;
}

/* CUDA-event timing of the closed-loop kernels on the mem's stream.  enable: on = 1.
 * get: tot[4] = summed ms of k2_build, k3_qp, k4_post over the steps since the last reset and the
 * number of steps; step_ms[cap] = per-step latency (build start -> post end). */
int eepacc_mem_enable_timing(int mem, int on)
{
    Mem* M = get_mem(mem);
    if (!M) return EEPACC_ERR_ARG;
    M->timing = on != 0;
    return EEPACC_OK;
}
int eepacc_mem_get_timing(int mem, double* tot, double* step_ms, long long cap)
{
    Mem* M = get_mem(mem);
    if (!M || !M->ready || !tot) return fail(EEPACC_ERR_ARG, "bad arguments");
    DeviceGuard dg_; CK((cudaError_t)dg_.enter(M->device));
    CK(cudaStreamSynchronize(M->stream));
    tot[0] = tot[1] = tot[2] = 0.0; tot[3] = (double)M->ev_steps;
    const long long sets = (long long)M->ev_steps * M->ev_groups;     // one set of events per (step, slice)
    for (long long k = 0; k < sets; ++k) {
        cudaEvent_t* ev = &M->evs[(size_t)k * 4];
        float a = 0, b = 0, c2 = 0;
        CK(cudaEventElapsedTime(&a, ev[0], ev[1]));
        CK(cudaEventElapsedTime(&b, ev[1], ev[2]));
        CK(cudaEventElapsedTime(&c2, ev[2], ev[3]));
        tot[0] += a; tot[1] += b; tot[2] += c2;
        if (step_ms && k < cap) step_ms[k] = (double)a + b + c2;
    }
    return EEPACC_OK;
}

/* eepacc_run pipelining: the batch is cut into `groups` contiguous slices (1..32), each advancing through
 * its MPC steps on its own stream.  0 = automatic (8 slices of >= 256 vehicles).  Results do not depend
 * on the slicing (vehicles never interact). */
/* eepacc_run mode: 0 (default) = one set of launches per MPC step (sliced, see eepacc_mem_set_groups);
 * 1 = steps >= 1 run in the fused closed-loop kernel, one launch per cycle, no synchronisation between vehicles.
 * Event timing (eepacc_mem_enable_timing) always uses the per-step launches. */
/* Closed-loop QP solver: EEPACC_SOLVER_QPOASES (default; K3/K3r, the reference's dense qpOASES path) or
 * EEPACC_SOLVER_OCP (K6, stage-structured interior point: the reference's solverToUse = 0 / 2 formulations). */
int eepacc_mem_set_solver(int mem, int solver)
{
    Mem* M = get_mem(mem);
    if (!M || !M->ready) return fail(EEPACC_ERR_ARG, "mem not set up");
    if (solver == EEPACC_SOLVER_QPOASES) { M->solver = solver; return EEPACC_OK; }
    if (solver != EEPACC_SOLVER_OCP) return fail(EEPACC_ERR_ARG, "unknown solver");
    if (M->ctrl != EEPACC_FB && M->ctrl != EEPACC_AB)
        return fail(EEPACC_ERR_ARG, "the stage-structured solver covers FBMPC and ABMPC (the BL/TV LPs have no unique interior solution)");
    DeviceGuard gd;
    CK((cudaError_t)gd.enter(M->device));
    const long long items = eep_ocp_work_items(M->ctrl, M->n, M->m);
    if (M->qz.ensure((size_t)M->B * (NH + 1 + M->n) * 8)) return EEPACC_ERR_MEM;
    if (M->ocp_work.ensure((size_t)M->B * (size_t)items * 8)) return EEPACC_ERR_MEM;
    M->solver = solver;
    M->fused = 0;               // the fused closed-loop kernel embeds the active-set solver
    // initial barrier parameter, measured on the bench batch: the convex ABMPC QPs want a well-centred
    // start (1e6: no iteration-limit failures in 13 936 QPs, 1e4: 109); on the non-convex FBMPC QPs the start decides
    // WHICH local minimiser is reached, and 1e4 is the value that reproduces the golden runs (qpOASES' minimiser)
    M->ocp_opts[0] = (M->ctrl == EEPACC_FB) ? 1e4 : 1e6;
    return EEPACC_OK;
}
int eepacc_mem_get_solver(int mem)
{
    Mem* M = get_mem(mem);
    return (M && M->ready) ? M->solver : -1;
}
/* opts[6] = mu0, eps_mu, eps_p, eps_d, eps_accept, maxit (NaN keeps a value) */
int eepacc_mem_set_ocp_options(int mem, const double* opts)
{
    Mem* M = get_mem(mem);
    if (!M || !M->ready || !opts) return fail(EEPACC_ERR_ARG, "mem not set up");
    for (int i = 0; i < 6; ++i) if (opts[i] == opts[i]) {
        if (!(opts[i] > 0.0)) return fail(EEPACC_ERR_ARG, "solver options must be positive");
        M->ocp_opts[i] = opts[i];
    }
    return EEPACC_OK;
}
int eepacc_mem_set_fused(int mem, int on)
{
    Mem* M = get_mem(mem);
    if (!M || !M->ready) return fail(EEPACC_ERR_ARG, "mem not set up");
    if (on && M->solver != EEPACC_SOLVER_QPOASES)
        return fail(EEPACC_ERR_ARG, "the fused closed-loop kernel embeds the active-set solver: select EEPACC_SOLVER_QPOASES first");
    M->fused = on ? 1 : 0;
    return EEPACC_OK;
}
/* Duration (ms, CUDA events on the launching stream) of the last fused closed-loop launch; out[1] = its grid,
 * out[2] = CTAs resident per SM.  Waits for the launch to finish. */
int eepacc_mem_get_cycle_timing(int mem, double* out)
{
    Mem* M = get_mem(mem);
    if (!M || !M->ready || !out) return fail(EEPACC_ERR_ARG, "mem not set up");
    if (!M->cyc_ev[1]) return fail(EEPACC_ERR_ARG, "no fused launch yet");
    CK(cudaEventSynchronize(M->cyc_ev[1]));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, M->cyc_ev[0], M->cyc_ev[1]));
    out[0] = ms; out[1] = M->cyc_grid; out[2] = M->cyc_per_sm;
    unsigned long long ph[4] = { 0, 0, 0, 0 };
    CK(cudaMemcpy(ph, M->cyc_counter.as<unsigned long long>() + 1, sizeof ph, cudaMemcpyDeviceToHost));
    for (int k = 0; k < 4; ++k) out[3 + k] = (double)ph[k];
    return EEPACC_OK;
}
int eepacc_mem_get_fused(int mem)
{
    Mem* M = get_mem(mem);
    return (M && M->ready) ? M->fused : -1;
}

int eepacc_mem_set_groups(int mem, int groups)
{
    Mem* M = get_mem(mem);
    if (!M || !M->ready) return fail(EEPACC_ERR_ARG, "mem not set up");
    if (groups < 0 || groups > MAX_GROUPS) return fail(EEPACC_ERR_ARG, "groups must be 0 (auto) .. 32");
    if (groups == 0) { groups = (int)(M->B / 256); if (groups > 16) groups = 16; if (groups < 1) groups = 1; }
    if (groups > M->B) groups = (int)M->B;
    M->n_groups = groups;
    return EEPACC_OK;
}
int eepacc_mem_get_groups(int mem)
{
    Mem* M = get_mem(mem);
    return (M && M->ready) ? M->n_groups : -1;
}

/* Developer aid: enable (cap > 0) / fetch the per-iteration trace of vehicle 0's last solve:
 * out[cap][4] = tau, ratio-test key, |W|, exchanges.  Returns 0 on success. */
int eepacc_mem_debug_trace(int mem, double* out, int cap)
{
    Mem* M = get_mem(mem);
    if (!M || !M->ready) return fail(EEPACC_ERR_ARG, "mem not set up");
    DeviceGuard dg_; CK((cudaError_t)dg_.enter(M->device));
    if (!M->dbg.p) {
        if (M->dbg.ensure((4096 * 4 + 32) * 8)) return EEPACC_ERR_MEM;
        CK(cudaMemset(M->dbg.p, 0, (4096 * 4 + 32) * 8));
        return EEPACC_OK;
    }
    if (out && cap > 0) {
        if (cap > 4096) cap = 4096;
        CK(cudaMemcpy(out, M->dbg.p, (size_t)cap * 4 * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemset(M->dbg.p, 0, 4096 * 4 * 8));
    }
    return EEPACC_OK;
}

/* Developer aid: copy one of the mem's per-step device buffers to the host (after eepacc_step):
 * what 0 sblk [B][EEP_SBLK], 1 g [B][n], 2 lba [B][m], 3 uba [B][m], 4 x [B][n], 5 status [B], 6 iter [B], 7 cost [B].
 * Returns the number of doubles written (<= cap) or a negative error. */
long long eepacc_mem_debug_fetch(int mem, int what, double* out, long long cap)
{
    Mem* M = get_mem(mem);
    if (!M || !M->ready || !out) { fail(EEPACC_ERR_ARG, "bad arguments"); return -1; }
    DeviceGuard dg_; if (dg_.enter(M->device) != 0) return -1;
    const long long B = M->B;
    const void* src = nullptr; long long cnt = 0;
    switch (what) {
    case 0: src = M->sblk.p; cnt = B * EEP_SBLK; break;
    case 1: src = M->g.p; cnt = B * M->n; break;
    case 2: src = M->lba.p; cnt = B * M->m; break;
    case 3: src = M->uba.p; cnt = B * M->m; break;
    case 4: src = M->xsol.p; cnt = B * M->n; break;
    case 5: src = M->status.p; cnt = B; break;
    case 6: src = M->iter.p; cnt = B; break;
    case 7: src = M->cost.p; cnt = B; break;
    case 9: src = M->lead.p; cnt = M->n_lead * B * 2; break;
    case 8: src = M->defer_cnt.p; cnt = 2; break;   // int[4]: vehicles deferred to the full-space kernel in the last step
    case 10: {
        // working set after the last solve, [B][m + n] as doubles: -1 held at its lower bound, +1 upper, 0 inactive
        // (rows 0..m-1, then the implicit variable bounds) -- the layout tests/ compare against (orc_vehicle_get_ws)
        const long long mt = M->m + M->n, n = M->n;
        if (cap < B * mt) { fail(EEPACC_ERR_ARG, "working-set fetch needs B*(m+n) doubles"); return -1; }
        if (cudaStreamSynchronize(M->stream) != cudaSuccess) return -1;
        std::vector<short> widx((size_t)(B * n)); std::vector<signed char> wside((size_t)(B * n));
        if (cudaMemcpy(widx.data(), M->widx.p, (size_t)(B * n) * 2, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        if (cudaMemcpy(wside.data(), M->wside.p, (size_t)(B * n), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        for (long long i = 0; i < B * mt; ++i) out[i] = 0.0;
        for (long long b = 0; b < B; ++b)
            for (long long i = 0; i < n; ++i) {
                const int c = widx[(size_t)(b * n + i)];
                if (c >= 0 && c < mt) out[b * mt + c] = (double)wside[(size_t)(b * n + i)];
            }
        return B * mt;
    }
    default: fail(EEPACC_ERR_ARG, "unknown buffer"); return -1;
    }
    if (cnt > cap) cnt = cap;
    if (cudaStreamSynchronize(M->stream) != cudaSuccess) return -1;
    if (cudaMemcpy(out, src, (size_t)cnt * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return cnt;
}

// ---------------- CasADi-style families ----------------
#define FAMILY_META(F, NIN, NOUT, INS, OUTS)                                                              \
    long long F##_n_in(void) { return NIN; }                                                              \
    long long F##_n_out(void) { return NOUT; }                                                            \
    const char* F##_name_in(long long i) { static const Port p[] = INS; return (i >= 0 && i < NIN) ? p[i].name : 0; }   \
    const char* F##_name_out(long long i) { static const Port p[] = OUTS; return (i >= 0 && i < NOUT) ? p[i].name : 0; } \
    const long long* F##_sparsity_in(long long i) { static const Port p[] = INS; if (i < 0 || i >= NIN) return 0;        \
        int r, c; resolve(p[i], F##_ctrl(), F##_m(), r, c); return dense_ccs(r, c); }                    \
    const long long* F##_sparsity_out(long long i) { static const Port p[] = OUTS; if (i < 0 || i >= NOUT) return 0;     \
        int r, c; resolve(p[i], F##_ctrl(), F##_m(), r, c); return dense_ccs(r, c); }                    \
    int F##_work(long long* sz_arg, long long* sz_res, long long* sz_iw, long long* sz_w) {              \
        if (sz_arg) *sz_arg = NIN; if (sz_res) *sz_res = NOUT; if (sz_iw) *sz_iw = 2; if (sz_w) *sz_w = 0; return 0; } \
    void F##_incref(void) { std::lock_guard<std::mutex> lk(g_mu); ++g_refcount; }                         \
    void F##_decref(void) { std::lock_guard<std::mutex> lk(g_mu); --g_refcount; }                         \
    int F##_checkout(void) { return checkout_impl(); }                                                    \
    void F##_release(int mem) { release_impl(mem); }                                                      \
    double F##_default_in(long long) { return 0.0; }

#define BR(...) { __VA_ARGS__ }
#define FIXED_CTRL(F, CTRL, MEXPR) static int F##_ctrl() { return CTRL; } static int F##_m() { return MEXPR; }

FIXED_CTRL(eepacc_f, g_last_ctrl, g_last_m)
FAMILY_META(eepacc_f, 2, 2, BR({"x", 4, 1}, {"u", 6, 1}), BR({"xdot", 4, 1}, {"L", 1, 1}))
FIXED_CTRL(eepacc_F, g_last_ctrl, g_last_m)
FAMILY_META(eepacc_F, 2, 2, BR({"x0", 4, 1}, {"p", 6, 1}), BR({"xf", 4, 1}, {"qf", 1, 1}))
FIXED_CTRL(eepacc_F_jac, g_last_ctrl, g_last_m)
FAMILY_META(eepacc_F_jac, 2, 3, BR({"x0", 4, 1}, {"p", 6, 1}), BR({"xf", 4, 1}, {"qf", 1, 1}, {"jac", 5, 10}))
FIXED_CTRL(eepacc_power5, g_last_ctrl, g_last_m)
FAMILY_META(eepacc_power5, 2, 1, BR({"Fm", 1, 1}, {"rpm", 1, 1}), BR({"P", 1, 1}))
FIXED_CTRL(eepacc_power2, g_last_ctrl, g_last_m)
FAMILY_META(eepacc_power2, 2, 1, BR({"Fm", 1, 1}, {"rpm", 1, 1}), BR({"P", 1, 1}))
FIXED_CTRL(eepacc_plant, g_last_ctrl, g_last_m)
FAMILY_META(eepacc_plant, 2, 1, BR({"x", 2, 1}, {"u", 2, 1}), BR({"x_next", 2, 1}))

#define QPB_INS BR({"x0", 2, 1}, {"a_minus1", 1, 1}, {"lead", 3, 1}, {"t0", 1, 1})
#define QPB_OUTS BR({"h", DIM_NN, 0}, {"g", DIM_N, 0}, {"a", DIM_MN, 0}, {"lba", DIM_M, 0}, {"uba", DIM_M, 0})
FIXED_CTRL(eepacc_qpbuild_ab, EEPACC_AB, (g_last_ctrl == EEPACC_AB ? g_last_m : 362))
FAMILY_META(eepacc_qpbuild_ab, 4, 5, QPB_INS, QPB_OUTS)
FIXED_CTRL(eepacc_qpbuild_fb, EEPACC_FB, 522)
FAMILY_META(eepacc_qpbuild_fb, 4, 5, QPB_INS, QPB_OUTS)
FIXED_CTRL(eepacc_qpbuild_bl, EEPACC_BL, 262)
FAMILY_META(eepacc_qpbuild_bl, 4, 5, QPB_INS, QPB_OUTS)
FIXED_CTRL(eepacc_qpbuild_tv, EEPACC_TV, 220)
FAMILY_META(eepacc_qpbuild_tv, 4, 5, QPB_INS, QPB_OUTS)

FIXED_CTRL(eepacc_qpsolve, g_last_ctrl, g_last_m)
FAMILY_META(eepacc_qpsolve, 10, 6,
            BR({"h", DIM_NN, 0}, {"g", DIM_N, 0}, {"a", DIM_MN, 0}, {"lba", DIM_M, 0}, {"uba", DIM_M, 0},
               {"lbx", DIM_N, 0}, {"ubx", DIM_N, 0}, {"x0", DIM_N, 0}, {"lam_x0", DIM_N, 0}, {"lam_a0", DIM_M, 0}),
            BR({"x", DIM_N, 0}, {"cost", 1, 1}, {"lam_a", DIM_M, 0}, {"lam_x", DIM_N, 0}, {"status", 1, 1}, {"iter", 1, 1}))

#define STEP_INS BR({"lead", 2, 1})
#define STEP_OUTS BR({"z1", 8, 1}, {"a", 1, 1}, {"status", 1, 1}, {"iter", 1, 1})
FIXED_CTRL(eepacc_step_ab, EEPACC_AB, 362)
FAMILY_META(eepacc_step_ab, 1, 4, STEP_INS, STEP_OUTS)
FIXED_CTRL(eepacc_step_fb, EEPACC_FB, 522)
FAMILY_META(eepacc_step_fb, 1, 4, STEP_INS, STEP_OUTS)
FIXED_CTRL(eepacc_step_bl, EEPACC_BL, 262)
FAMILY_META(eepacc_step_bl, 1, 4, STEP_INS, STEP_OUTS)
FIXED_CTRL(eepacc_step_tv, EEPACC_TV, 220)
FAMILY_META(eepacc_step_tv, 1, 4, STEP_INS, STEP_OUTS)

FIXED_CTRL(eepacc_run, g_last_ctrl, g_last_m)
FAMILY_META(eepacc_run, 1, 2, BR({"n_steps", 1, 1}), BR({"log", EEPACC_LOG_W, 1}, {"kpi", EEPACC_N_KPI, 1}))

// ---------------- evaluation ----------------
static int eval_k1(int which, const double** arg, double** res, long long* iw, int mem)
{
    CallCtx c;
    int rc = begin_call(mem, iw, c);
    if (rc) return rc;
    if (!arg || !res) return fail(EEPACC_ERR_ARG, "arg/res is NULL");
    Mem* M = c.M; const long long B = c.B;
    if (B == 0) return EEPACC_OK;
    const int nx = (which >= 3) ? 1 : 4, nuu = (which >= 3) ? 1 : 6;
    const double *dx, *du;
    if ((rc = stage_in(M, 0, arg[0], (size_t)B * nx, c.dev, &dx))) return rc;
    if ((rc = stage_in(M, 1, arg[1], (size_t)B * nuu, c.dev, &du))) return rc;
    if (!dx || !du) return fail(EEPACC_ERR_ARG, "missing input");
    double *o0, *o1 = nullptr, *o2 = nullptr;
    const size_t n0 = (which >= 3) ? 1 : 4;
    if ((rc = stage_out(M, 0, res[0], (size_t)B * n0, c.dev, &o0))) return rc;
    if (which <= 2 && (rc = stage_out(M, 1, res[1], (size_t)B, c.dev, &o1))) return rc;
    if (which == 2 && (rc = stage_out(M, 2, res[2], (size_t)B * 50, c.dev, &o2))) return rc;
    if (!o0 || (which <= 2 && !o1) || (which == 2 && !o2)) return fail(EEPACC_ERR_ARG, "missing output buffer");
    switch (which) {
    case 0: eep_launch_f(&M->p, B, dx, du, o0, o1, M->stream); break;
    case 1: eep_launch_F(&M->p, B, dx, du, o0, o1, M->stream); break;
    case 2: eep_launch_Fjac(&M->p, B, dx, du, o0, o1, o2, M->stream); break;
    case 3: eep_launch_power(&M->p, B, 1, dx, du, o0, M->stream); break;
    default: eep_launch_power(&M->p, B, 0, dx, du, o0, M->stream); break;
    }
    M->launches += 1;
    if ((rc = copy_out(M, 0, res[0], (size_t)B * n0, c.dev))) return rc;
    if (which <= 2 && (rc = copy_out(M, 1, res[1], (size_t)B, c.dev))) return rc;
    if (which == 2 && (rc = copy_out(M, 2, res[2], (size_t)B * 50, c.dev))) return rc;
    return end_call(c);
}

int eepacc_f(const double** arg, double** res, long long* iw, double*, int mem) { return eval_k1(0, arg, res, iw, mem); }
int eepacc_F(const double** arg, double** res, long long* iw, double*, int mem) { return eval_k1(1, arg, res, iw, mem); }
int eepacc_F_jac(const double** arg, double** res, long long* iw, double*, int mem) { return eval_k1(2, arg, res, iw, mem); }
int eepacc_power5(const double** arg, double** res, long long* iw, double*, int mem) { return eval_k1(3, arg, res, iw, mem); }
int eepacc_power2(const double** arg, double** res, long long* iw, double*, int mem) { return eval_k1(4, arg, res, iw, mem); }

int eepacc_plant(const double** arg, double** res, long long* iw, double*, int mem)
{
    CallCtx c;
    int rc = begin_call(mem, iw, c);
    if (rc) return rc;
    if (!arg || !res) return fail(EEPACC_ERR_ARG, "arg/res is NULL");
    Mem* M = c.M; const long long B = c.B;
    if (B == 0) return EEPACC_OK;
    if (B > M->B) return fail(EEPACC_ERR_ARG, "batch exceeds the routes bound to mem");
    const double *dx, *du; double* o;
    if ((rc = stage_in(M, 0, arg[0], (size_t)B * 2, c.dev, &dx))) return rc;
    if ((rc = stage_in(M, 1, arg[1], (size_t)B * 2, c.dev, &du))) return rc;
    if ((rc = stage_out(M, 0, res[0], (size_t)B * 2, c.dev, &o))) return rc;
    if (!dx || !du || !o) return fail(EEPACC_ERR_ARG, "missing buffer");
    eep_launch_plant(&M->p, B, M->routes.as<eepacc_route>(), 1, dx, du, o, M->stream);
    M->launches += 1;
    if ((rc = copy_out(M, 0, res[0], (size_t)B * 2, c.dev))) return rc;
    return end_call(c);
}

static int eval_qpbuild(int ctrl, const double** arg, double** res, long long* iw, int mem)
{
    CallCtx c;
    int rc = begin_call(mem, iw, c);
    if (rc) return rc;
    Mem* M = c.M;
    if (M->ctrl != ctrl) return fail(EEPACC_ERR_ARG, "mem was set up for a different controller");
    if (c.B != M->B) return fail(EEPACC_ERR_ARG, "batch must equal the batch bound to mem");
    if (!arg || !res) return fail(EEPACC_ERR_ARG, "arg/res is NULL");
    const size_t B = (size_t)M->B, n = M->n, m = M->m;
    eep_build_io io = make_build_io(M, 1);
    if ((rc = stage_in(M, 0, arg[0], B * 2, c.dev, &io.in_x0))) return rc;
    if ((rc = stage_in(M, 1, arg[1], B, c.dev, &io.in_am1))) return rc;
    if ((rc = stage_in(M, 2, arg[2], B * 3, c.dev, &io.in_lead))) return rc;
    if ((rc = stage_in(M, 3, arg[3], B, c.dev, &io.in_t0))) return rc;
    if (!io.in_x0 || !io.in_am1 || !io.in_lead || !io.in_t0) return fail(EEPACC_ERR_ARG, "missing input");
    const size_t cnt[5] = { B * n * n, B * n, B * m * n, B * m, B * m };
    double* o[5];
    for (int k = 0; k < 5; ++k) {
        if ((rc = stage_out(M, k, res[k], cnt[k], c.dev, &o[k]))) return rc;
        if (!o[k]) return fail(EEPACC_ERR_ARG, "missing output buffer");
    }
    io.H = o[0]; io.g = o[1]; io.A = o[2]; io.lba = o[3]; io.uba = o[4];
    io.sblk = nullptr; io.dyn = nullptr;
    CK(cudaMemsetAsync(io.A, 0, cnt[2] * 8, M->stream));
    rc = eep_launch_build(&M->p, &io, M->stream);
    if (rc) return fail(EEPACC_ERR_CUDA, "k2_build launch", (cudaError_t)rc);
    M->launches += 1;
    for (int k = 0; k < 5; ++k) if ((rc = copy_out(M, k, res[k], cnt[k], c.dev))) return rc;
    return end_call(c);
}
int eepacc_qpbuild_ab(const double** a, double** r, long long* iw, double*, int mem) { return eval_qpbuild(EEPACC_AB, a, r, iw, mem); }
int eepacc_qpbuild_fb(const double** a, double** r, long long* iw, double*, int mem) { return eval_qpbuild(EEPACC_FB, a, r, iw, mem); }
int eepacc_qpbuild_bl(const double** a, double** r, long long* iw, double*, int mem) { return eval_qpbuild(EEPACC_BL, a, r, iw, mem); }
int eepacc_qpbuild_tv(const double** a, double** r, long long* iw, double*, int mem) { return eval_qpbuild(EEPACC_TV, a, r, iw, mem); }

int eepacc_qpsolve(const double** arg, double** res, long long* iw, double*, int mem)
{
    CallCtx c;
    int rc = begin_call(mem, iw, c);
    if (rc) return rc;
    Mem* M = c.M;
    if (c.B != M->B) return fail(EEPACC_ERR_ARG, "batch must equal the batch bound to mem (warm-start state is per vehicle)");
    if (!arg || !res) return fail(EEPACC_ERR_ARG, "arg/res is NULL");
    const size_t B = (size_t)M->B, n = M->n, m = M->m;
    eep_qp_io io = make_qp_io(M);
    const size_t icnt[7] = { B * n * n, B * n, B * m * n, B * m, B * m, B * n, B * n };
    const double* in[7];
    for (int k = 0; k < 7; ++k) if ((rc = stage_in(M, k, arg[k], icnt[k], c.dev, &in[k]))) return rc;
    for (int k = 0; k < 5; ++k) if (!in[k]) return fail(EEPACC_ERR_ARG, "h, g, a, lba, uba are required");
    io.H = in[0]; io.g = in[1]; io.A = in[2]; io.lba = in[3]; io.uba = in[4]; io.lbx = in[5]; io.ubx = in[6];
    io.sblk = nullptr;          // dense conic operands: the generic operators
    /* arg[7..9] (x0, lam_x0, lam_a0) are accepted for scheme compatibility; the hot start lives in mem,
       exactly like the plugin's implicit qpOASES hot start. */
    const size_t ocnt[6] = { B * n, B, B * m, B * n, B, B };
    double* o[6];
    for (int k = 0; k < 6; ++k) if ((rc = stage_out(M, k, res[k], ocnt[k], c.dev, &o[k]))) return rc;
    if (!o[0]) return fail(EEPACC_ERR_ARG, "output x is required");
    io.x = o[0];
    if (o[1]) io.cost = o[1];
    io.lam_a = o[2]; io.lam_x = o[3];
    if (o[4]) io.status = o[4];
    if (o[5]) io.iter = o[5];
    rc = eep_launch_qp(&M->p, &io, M->stream);
    if (rc) return fail(EEPACC_ERR_CUDA, "k3_qp launch", (cudaError_t)rc);
    M->launches += 1;
    for (int k = 0; k < 6; ++k) if ((rc = copy_out(M, k, res[k], ocnt[k], c.dev))) return rc;
    return end_call(c);
}

static int eval_step(int ctrl, const double** arg, double** res, long long* iw, int mem)
{
    CallCtx c;
    int rc = begin_call(mem, iw, c);
    if (rc) return rc;
    Mem* M = c.M;
    if (M->ctrl != ctrl) return fail(EEPACC_ERR_ARG, "mem was set up for a different controller");
    if (c.B != M->B) return fail(EEPACC_ERR_ARG, "batch must equal the batch bound to mem");
    if (!res) return fail(EEPACC_ERR_ARG, "res is NULL");
    const size_t B = (size_t)M->B;
    const double* lead = nullptr;
    if (ctrl != EEPACC_TV) {
        if (!arg || !arg[0]) return fail(EEPACC_ERR_ARG, "lead input is required");
        if ((rc = stage_in(M, 0, arg[0], B * 2, c.dev, &lead))) return rc;
    }
    double* log = M->logrow.as<double>();
    if ((rc = device_step(M, lead, log))) return rc;
    // unpack the log row into the four outputs (device-side strided copies)
    const size_t W = EEPACC_LOG_W * 8;
    const int off[4] = { 0, 8, 9, 10 }, wid[4] = { 8, 1, 1, 1 };
    for (int k = 0; k < 4; ++k) {
        if (!res[k]) continue;
        CK(cudaMemcpy2DAsync(res[k], wid[k] * 8, log + off[k], W, wid[k] * 8, B,
                             c.dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, M->stream));
    }
    return end_call(c);
}
int eepacc_step_ab(const double** a, double** r, long long* iw, double*, int mem) { return eval_step(EEPACC_AB, a, r, iw, mem); }
int eepacc_step_fb(const double** a, double** r, long long* iw, double*, int mem) { return eval_step(EEPACC_FB, a, r, iw, mem); }
int eepacc_step_bl(const double** a, double** r, long long* iw, double*, int mem) { return eval_step(EEPACC_BL, a, r, iw, mem); }
int eepacc_step_tv(const double** a, double** r, long long* iw, double*, int mem) { return eval_step(EEPACC_TV, a, r, iw, mem); }

int eepacc_run(const double** arg, double** res, long long* iw, double*, int mem)
{
    CallCtx c;
    int rc = begin_call(mem, iw, c);
    if (rc) return rc;
    Mem* M = c.M;
    if (!arg || !arg[0] || !res) return fail(EEPACC_ERR_ARG, "n_steps input is required (host scalar)");
    const long long n_steps = (long long)arg[0][0];
    if (n_steps <= 0) return fail(EEPACC_ERR_ARG, "n_steps must be positive");
    if (c.B != M->B) return fail(EEPACC_ERR_ARG, "iw[0] (batch) differs from the batch the mem was set up for");
    if (M->ctrl != EEPACC_TV && (M->n_lead < n_steps || !M->lead.p)) return fail(EEPACC_ERR_ARG, "lead trace shorter than n_steps (eepacc_mem_set_lead)");
    const size_t B = (size_t)M->B;
    if ((rc = reset_state(M))) return rc;
    double* log_dev = nullptr;
    if (res[0]) {
        if (c.dev) log_dev = res[0];
        else { if (M->out[0].ensure((size_t)n_steps * B * EEPACC_LOG_W * 8)) return EEPACC_ERR_MEM; log_dev = M->out[0].as<double>(); }
    }
    const int G = M->n_groups;
    if (M->fused && M->qp_mode == 0 && !M->timing) {
        // step 0 (cold start: every vehicle goes to the full-space kernel, the pinned reference path) through the
        // per-step launches, everything after it in ONE launch with no synchronisation between vehicles
        {
            const double* lead = (M->ctrl == EEPACC_TV) ? nullptr : M->lead.as<double>();
            if ((rc = device_step(M, lead, log_dev))) return rc;
        }
        if (n_steps > 1) {
            eep_cycle_io cy;
            memset(&cy, 0, sizeof cy);
            cy.b = make_build_io(M, 0);
            cy.q = make_qp_io(M);
            cy.q.fc = M->qp_fc;
            cy.q.rowtab = M->rowtab.as<unsigned long long>(); cy.q.rowc = M->rowc.as<double>();
            cy.q.hscr = M->hscr.as<double>();
            cy.p = make_post_io(M, 0, nullptr);
            cy.lead_all = (M->ctrl == EEPACC_TV) ? nullptr : M->lead.as<double>();
            cy.lead_stride = (long long)B * 2;
            cy.log_all = log_dev; cy.log_stride = (long long)B * EEPACC_LOG_W;
            cy.k_begin = 1; cy.k_end = (int)n_steps; cy.t0_begin = M->t0;
            cy.fc_big = M->n;
            cy.big_bytes = (eep_qpr_smem_bytes(M->n, M->m, M->n) + 255) & ~(size_t)255;
            const int per_sm = eep_cycle_max_ctas_per_sm(&cy);
            if (per_sm < 1) return fail(EEPACC_ERR_CUDA, "fused closed-loop kernel does not fit an SM");
            int sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, M->device);
            long long grid = (long long)per_sm * sms;
            if (const char* e = getenv("EEPACC_CYCLE_GRID")) { long long v = atoll(e); if (v >= 1 && v < grid) grid = v; }   // developer knob
            if (grid > M->B) grid = M->B;
            if (M->cyc_counter.ensure(8 * 8) || M->cyc_big.ensure((size_t)grid * cy.big_bytes)) return EEPACC_ERR_MEM;
            cy.work_counter = M->cyc_counter.as<unsigned long long>();
            cy.big = M->cyc_big.as<unsigned char>();
            cy.phase = cy.work_counter + 1;
            CK(cudaMemsetAsync(cy.work_counter, 0, 8 * 8, M->stream));
            for (int e = 0; e < 2; ++e) if (!M->cyc_ev[e]) CK(cudaEventCreate(&M->cyc_ev[e]));
            CK(cudaEventRecord(M->cyc_ev[0], M->stream));
            rc = eep_launch_cycle(&M->p, &cy, (int)grid, M->stream);
            CK(cudaEventRecord(M->cyc_ev[1], M->stream));
            M->cyc_grid = (int)grid; M->cyc_per_sm = per_sm;
            if (rc) return fail(EEPACC_ERR_CUDA, "k5_cycle launch", (cudaError_t)rc);
            M->launches += 1;
            M->kk = (int)n_steps; M->t0 += (double)(n_steps - 1) * M->p.Ts;
        }
    } else if (G <= 1) {
        for (long long k = 0; k < n_steps; ++k) {
            const double* lead = (M->ctrl == EEPACC_TV) ? nullptr : M->lead.as<double>() + (size_t)k * B * 2;
            double* lg = log_dev ? log_dev + (size_t)k * B * EEPACC_LOG_W : nullptr;
            if ((rc = device_step(M, lead, lg))) return rc;
        }
    } else {
        // slices on their own streams: fork from the mem's stream, join back before the KPI kernel
        if (!M->gfork) CK(cudaEventCreateWithFlags(&M->gfork, cudaEventDisableTiming));
        for (int g = 0; g < G; ++g) {
            if (!M->gstream[g]) CK(cudaStreamCreateWithFlags(&M->gstream[g], cudaStreamNonBlocking));
            if (!M->gjoin[g]) CK(cudaEventCreateWithFlags(&M->gjoin[g], cudaEventDisableTiming));
        }
        if (M->timing) {
            if (timing_events(M, (size_t)n_steps * G * 4)) return EEPACC_ERR_CUDA;
            M->ev_steps = (int)n_steps; M->ev_groups = G;
        }
        CK(cudaEventRecord(M->gfork, M->stream));
        for (int g = 0; g < G; ++g) CK(cudaStreamWaitEvent(M->gstream[g], M->gfork, 0));
        double t0 = M->t0;
        for (long long k = 0; k < n_steps && !rc; ++k) {
            const double* lead = (M->ctrl == EEPACC_TV) ? nullptr : M->lead.as<double>() + (size_t)k * B * 2;
            double* lg = log_dev ? log_dev + (size_t)k * B * EEPACC_LOG_W : nullptr;
            for (int g = 0; g < G && !rc; ++g) {
                const size_t v0 = B * (size_t)g / G, v1 = B * (size_t)(g + 1) / G;
                cudaEvent_t* ev = M->timing ? &M->evs[((size_t)k * G + g) * 4] : nullptr;
                rc = device_step_slice(M, g, v0, (long long)(v1 - v0), M->gstream[g], (int)k, t0, lead, lg, ev);
            }
            t0 += M->p.Ts;
        }
        if (rc) {
            // a launch failed with slices in flight on their own (non-blocking) streams: drain them before
            // handing the failure back, so the caller may free or reuse its buffers
            for (int g = 0; g < G; ++g) cudaStreamSynchronize(M->gstream[g]);
            return rc;
        }
        for (int g = 0; g < G; ++g) {
            CK(cudaEventRecord(M->gjoin[g], M->gstream[g]));
            CK(cudaStreamWaitEvent(M->stream, M->gjoin[g], 0));
        }
        M->kk = (int)n_steps; M->t0 = t0;
    }
    rc = eep_launch_kpi_final(&M->p, M->B, (int)n_steps, M->kpi_acc.as<double>(), M->kpi.as<double>(), M->stream);
    if (rc) return fail(EEPACC_ERR_CUDA, "kpi launch", (cudaError_t)rc);
    M->launches += 1;
    if (res[0] && !c.dev) CK(cudaMemcpyAsync(res[0], log_dev, (size_t)n_steps * B * EEPACC_LOG_W * 8, cudaMemcpyDeviceToHost, M->stream));
    if (res[1]) CK(cudaMemcpyAsync(res[1], M->kpi.p, B * EEPACC_N_KPI * 8, c.dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, M->stream));
    return end_call(c);
}

}  // extern "C"
