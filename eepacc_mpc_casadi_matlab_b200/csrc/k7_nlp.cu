// k7_nlp.cu -- K7 launch + the C-ABI of include/eepacc_nlp.h (the batched NLP benchmark).
//
// One scenario per thread (k7_nlp.cuh).  The per-scenario workspace (K7_STAGE doubles per interval, 1.5 MB for the
// 870-interval cycle) lives in HBM as [interval][field][scenario]: the 32 lanes of a warp walk their scenarios'
// stages in lock step as long as their iteration counts agree, so every workspace access of a warp is one coalesced
// 256-byte line.  Blocks are 32 threads (one warp) so that a batch of B scenarios spreads over B / 32 SMs' worth of
// schedulers before any SM holds two warps: the solve is a long dependent chain per scenario (latency-bound, like K6).
// No CPU fallback: set-up fails with EEPACC_ERR_NO_DEVICE when there is no usable CUDA device.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <mutex>
#include "k7_nlp.cuh"

namespace {

struct NlpMem {
    bool used = false, ready = false;
    int device = -1, N = 0;
    long long B = 0;
    eepacc_params* p = nullptr;            // device copies
    eepacc_route* routes = nullptr;
    eepacc_nlp_tables* tabs = nullptr;
    double* ws = nullptr;                  // [N + 1][K7_STAGE][B]
    double *d_x0 = nullptr, *d_p = nullptr, *d_x = nullptr, *d_f = nullptr, *d_stats = nullptr;   // staging for host callers
    eepacc_nlp_opts opts;
    cudaEvent_t ev[2] = { nullptr, nullptr };
    double last_ms = 0.0;
};

constexpr int K7_MAX_MEM = 16;
NlpMem g_nlp[K7_MAX_MEM];
std::mutex g_nlp_mu;
thread_local char g_nlp_err[256] = "";

int nlp_fail(int code, const char* msg)
{
    snprintf(g_nlp_err, sizeof g_nlp_err, "%s", msg);
    return code;
}
int nlp_cuda_fail(cudaError_t e, const char* what)
{
    snprintf(g_nlp_err, sizeof g_nlp_err, "%s: %s", what, cudaGetErrorString(e));
    return EEPACC_ERR_CUDA;
}
#define K7_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return nlp_cuda_fail(e_, #call); } while (0)

// restores the caller's current device on every exit path (set-up and eval switch to the mem's device)
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

NlpMem* nlp_get(int mem) { return (mem >= 0 && mem < K7_MAX_MEM && g_nlp[mem].used) ? &g_nlp[mem] : nullptr; }

void nlp_free_all(NlpMem* M)
{
    if (M->device >= 0) {
        int prev = -1;
        cudaGetDevice(&prev);
        cudaSetDevice(M->device);
        cudaFree(M->p); cudaFree(M->routes); cudaFree(M->tabs); cudaFree(M->ws);
        cudaFree(M->d_x0); cudaFree(M->d_p); cudaFree(M->d_x); cudaFree(M->d_f); cudaFree(M->d_stats);
        for (auto& e : M->ev) if (e) { cudaEventDestroy(e); e = nullptr; }
        if (prev >= 0) cudaSetDevice(prev);
    }
    M->p = nullptr; M->routes = nullptr; M->tabs = nullptr; M->ws = nullptr;
    M->d_x0 = M->d_p = M->d_x = M->d_f = M->d_stats = nullptr;
    M->ready = false; M->B = 0; M->N = 0; M->device = -1; M->last_ms = 0.0;
}

struct K7Args {
    const eepacc_params* p;
    const eepacc_route* routes;
    const eepacc_nlp_tables* tabs;
    const double* x0;        // [B][nz] or null
    const double* s_tv;      // [B][N + 1]
    double* ws;
    double* x;               // [B][nz]
    double* f;               // [B]
    double* stats;           // [B][EEPACC_NLP_N_STATS]
    eepacc_nlp_opts opts;
    long long B;
    int N;
};

}  // namespace

__global__ void __launch_bounds__(32) k7_nlp_kernel(const __grid_constant__ K7Args a)
{
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const long long nz = 4 + 10LL * a.N;
    K7Ctx c;
    k7_ctx_init(c, a.p, a.routes + b, a.tabs + b, a.s_tv + b * (a.N + 1), a.N, a.opts.obj_scale);
    K7Work w;
    w.ws = a.ws + b;
    w.stride = a.B;
    if (a.x0) k7_read_guess(c, w, a.x0 + b * nz); else k7_start(c, w);
    K7Result r;
    r.status = EEPACC_NLP_MAXITER; r.iters = 0; r.backtracks = 0;
    r.f = r.kkt = r.infeas = r.compl_ = r.mu = r.delta = 0.0;
    k7_solve(c, w, a.opts, r);
    k7_write_z(c, w, a.x + b * nz);
    a.f[b] = r.f;
    double* st = a.stats + b * EEPACC_NLP_N_STATS;
    st[0] = r.status; st[1] = r.iters; st[2] = r.kkt; st[3] = r.infeas; st[4] = r.compl_;
    st[5] = r.mu; st[6] = r.delta; st[7] = r.backtracks;
}

extern "C" {

void eepacc_nlp_default_opts(eepacc_nlp_opts* o)
{
    o->tol = 1e-8;            // ipopt default tol; RunOpt_NLP.m sets acceptable_tol = 1e-8 as well (:253)
    o->mu0 = 1.0;
    o->obj_scale = 1e-4;
    o->mu_min = 1e-9;
    o->acceptable_kkt = 1.0;
    o->max_iter = 500;        // OPTsettings.NLPmaxIter is 5e3 for ipopt started from zero (Settings.m:91)
    o->pad_ = 0;
}

int eepacc_nlp_checkout(void)
{
    std::lock_guard<std::mutex> lk(g_nlp_mu);
    for (int i = 0; i < K7_MAX_MEM; ++i)
        if (!g_nlp[i].used) {
            g_nlp[i] = NlpMem();
            g_nlp[i].used = true;
            eepacc_nlp_default_opts(&g_nlp[i].opts);
            return i;
        }
    nlp_fail(EEPACC_ERR_MEM, "no free NLP memory slot");
    return -1;
}

void eepacc_nlp_release(int mem)
{
    std::lock_guard<std::mutex> lk(g_nlp_mu);
    NlpMem* M = nlp_get(mem);
    if (!M) return;
    nlp_free_all(M);
    M->used = false;
}

long long eepacc_nlp_nz(int n_steps) { return 4 + 10LL * n_steps; }
long long eepacc_nlp_workspace_bytes(long long B, int n_steps) { return (long long)sizeof(double) * K7_STAGE * (n_steps + 1LL) * B; }
long long eepacc_nlp_sizeof_tables(void) { return (long long)sizeof(eepacc_nlp_tables); }
const char* eepacc_nlp_last_error(void) { return g_nlp_err; }
long long eepacc_nlp_n_in(void) { return 2; }
long long eepacc_nlp_n_out(void) { return 3; }
const char* eepacc_nlp_name_in(long long i) { return i == 0 ? "x0" : i == 1 ? "p" : nullptr; }       // nlpsol scheme
const char* eepacc_nlp_name_out(long long i) { return i == 0 ? "x" : i == 1 ? "f" : i == 2 ? "stats" : nullptr; }
int eepacc_nlp_work(long long* sz_arg, long long* sz_res, long long* sz_iw, long long* sz_w)
{
    if (sz_arg) *sz_arg = 2;
    if (sz_res) *sz_res = 3;
    if (sz_iw) *sz_iw = 2;
    if (sz_w) *sz_w = 0;       // the workspace belongs to the mem object (device memory)
    return 0;
}

int eepacc_nlp_set_opts(int mem, const eepacc_nlp_opts* o)
{
    NlpMem* M = nlp_get(mem);
    if (!M || !o) return nlp_fail(EEPACC_ERR_ARG, "bad mem or options");
    if (!(o->tol > 0) || !(o->mu0 > 0) || !(o->obj_scale > 0) || !(o->mu_min > 0) || !(o->acceptable_kkt > 0) || o->max_iter < 1 || o->max_iter > 100000)
        return nlp_fail(EEPACC_ERR_ARG, "options out of range");
    M->opts = *o;
    return EEPACC_OK;
}

int eepacc_nlp_setup(int mem, const eepacc_params* p, const eepacc_route* routes, const eepacc_nlp_tables* tables,
                     long long B, int n_steps, int device)
{
    NlpMem* M = nlp_get(mem);
    if (!M) return nlp_fail(EEPACC_ERR_ARG, "bad mem");
    if (!p || !routes || !tables || B < 1 || n_steps < 1) return nlp_fail(EEPACC_ERR_ARG, "bad set-up arguments");
    for (long long b = 0; b < B; ++b) {
        if (tables[b].n_vi < 1 || tables[b].n_vi > EEPACC_NLP_MAX_VI || tables[b].n_stop < 1 ||
            tables[b].n_stop > EEPACC_NLP_MAX_STOPV || routes[b].n_tl < 0 || routes[b].n_tl > EEPACC_MAX_TL)
            return nlp_fail(EEPACC_ERR_ARG, "a scenario's lookup table has a bad point count");
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1 || device < 0 || device >= ndev)
        return nlp_fail(EEPACC_ERR_NO_DEVICE, "no usable CUDA device: the NLP benchmark has no CPU fallback");
    nlp_free_all(M);
    DeviceGuard guard;
    K7_CUDA(cudaSetDevice(device));
    M->device = device; M->B = B; M->N = n_steps;
    const long long nz = eepacc_nlp_nz(n_steps);
    K7_CUDA(cudaMalloc(&M->p, sizeof(eepacc_params)));
    K7_CUDA(cudaMalloc(&M->routes, sizeof(eepacc_route) * B));
    K7_CUDA(cudaMalloc(&M->tabs, sizeof(eepacc_nlp_tables) * B));
    K7_CUDA(cudaMalloc(&M->ws, (size_t)eepacc_nlp_workspace_bytes(B, n_steps)));
    K7_CUDA(cudaMalloc(&M->d_x0, sizeof(double) * nz * B));
    K7_CUDA(cudaMalloc(&M->d_p, sizeof(double) * (n_steps + 1LL) * B));
    K7_CUDA(cudaMalloc(&M->d_x, sizeof(double) * nz * B));
    K7_CUDA(cudaMalloc(&M->d_f, sizeof(double) * B));
    K7_CUDA(cudaMalloc(&M->d_stats, sizeof(double) * EEPACC_NLP_N_STATS * B));
    K7_CUDA(cudaMemcpy(M->p, p, sizeof(eepacc_params), cudaMemcpyHostToDevice));
    K7_CUDA(cudaMemcpy(M->routes, routes, sizeof(eepacc_route) * B, cudaMemcpyHostToDevice));
    K7_CUDA(cudaMemcpy(M->tabs, tables, sizeof(eepacc_nlp_tables) * B, cudaMemcpyHostToDevice));
    K7_CUDA(cudaMemset(M->ws, 0, (size_t)eepacc_nlp_workspace_bytes(B, n_steps)));
    K7_CUDA(cudaEventCreate(&M->ev[0]));
    K7_CUDA(cudaEventCreate(&M->ev[1]));
    M->ready = true;
    return EEPACC_OK;
}

int eepacc_nlp(const double** arg, double** res, long long* iw, double* w, int mem)
{
    (void)w;
    NlpMem* M = nlp_get(mem);
    if (!M || !M->ready) return nlp_fail(EEPACC_ERR_ARG, "mem not set up");
    if (!arg || !res || !iw) return nlp_fail(EEPACC_ERR_ARG, "null arg / res / iw");
    if (iw[0] != M->B) return nlp_fail(EEPACC_ERR_ARG, "iw[0] differs from the set-up batch size");
    if (!arg[1] || !res[0] || !res[1] || !res[2]) return nlp_fail(EEPACC_ERR_ARG, "p, x, f and stats are required");
    const bool dev = (iw[1] & EEPACC_FLAG_DEVICE_PTR) != 0;
    const long long B = M->B, nz = eepacc_nlp_nz(M->N);
    DeviceGuard guard;
    K7_CUDA(cudaSetDevice(M->device));
    K7Args a;
    a.p = M->p; a.routes = M->routes; a.tabs = M->tabs; a.ws = M->ws;
    a.opts = M->opts; a.B = B; a.N = M->N;
    if (dev) {
        a.x0 = arg[0]; a.s_tv = arg[1]; a.x = res[0]; a.f = res[1]; a.stats = res[2];
    } else {
        if (arg[0]) K7_CUDA(cudaMemcpy(M->d_x0, arg[0], sizeof(double) * nz * B, cudaMemcpyHostToDevice));
        K7_CUDA(cudaMemcpy(M->d_p, arg[1], sizeof(double) * (M->N + 1LL) * B, cudaMemcpyHostToDevice));
        a.x0 = arg[0] ? M->d_x0 : nullptr; a.s_tv = M->d_p; a.x = M->d_x; a.f = M->d_f; a.stats = M->d_stats;
    }
    const unsigned grid = (unsigned)((B + 31) / 32);
    K7_CUDA(cudaEventRecord(M->ev[0], 0));
    k7_nlp_kernel<<<grid, 32>>>(a);
    K7_CUDA(cudaGetLastError());
    K7_CUDA(cudaEventRecord(M->ev[1], 0));
    K7_CUDA(cudaEventSynchronize(M->ev[1]));
    float ms = 0.f;
    K7_CUDA(cudaEventElapsedTime(&ms, M->ev[0], M->ev[1]));
    M->last_ms = ms;
    if (!dev) {
        K7_CUDA(cudaMemcpy(res[0], M->d_x, sizeof(double) * nz * B, cudaMemcpyDeviceToHost));
        K7_CUDA(cudaMemcpy(res[1], M->d_f, sizeof(double) * B, cudaMemcpyDeviceToHost));
        K7_CUDA(cudaMemcpy(res[2], M->d_stats, sizeof(double) * EEPACC_NLP_N_STATS * B, cudaMemcpyDeviceToHost));
    }
    return EEPACC_OK;
}

double eepacc_nlp_last_ms(int mem)
{
    NlpMem* M = nlp_get(mem);
    return M ? M->last_ms : -1.0;
}

}  // extern "C"
