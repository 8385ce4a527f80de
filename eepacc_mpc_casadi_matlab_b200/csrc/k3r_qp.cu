// k3r_qp.cu -- K3 (closed-loop path): batched, warm-started parametric active-set QP solver in the
// REDUCED space of the free variables, one vehicle per CTA, several CTAs per SM.
//
// Stands where the reference calls the CasADi conic plugin 'qpoases' once per MPC step
//     sol = QPsolver('h',H,'g',c,'a',G,'lbx',..,'ubx',..,'lba',g_lb,'uba',g_ub)
// (MC/RunOpt_ABMPC.m:252, RunOpt_FBMPC.m:278, RunOpt_BLMPC.m:230, RunOpt_TVMPC.m:208).
// Same algorithm, tolerances and tie-break rules as k3_qp.cu / oracle/qp_oracle.c (homotopy, far
// bounds, exchange steps, anti-cycling); what changes is the factorisation it is carried out on.
//
// Observation (measured on the perturbed urban/highway batches): of the ~90-105 active rows of a
// 100/120-variable QP, ~3/4 are UNIT rows (slack >= 0, Fb <= 0, far bounds: one entry on one
// variable).  A variable pinned by an active unit row is "fixed"; only nF ~ 40 variables are free,
// the null space has nZ <= 20 columns.  As qpOASES does for its bounds, the factorisation
//     A_gF Qf = [0 T],   Z' H_FF Z = R'R,     Qf = [Z Y]  (nF x nF)
// only covers the free variables F and the active GENERAL rows g; the fixed block is implicit:
//     full Q = diag(Qf, I_X),   full T = [T  A_gX; 0  diag(cv_X)].
// Unit rows that become active during a solve are handled as general rows (no shrinking); the
// compaction happens at the next MPC step's warm start, where everything is rebuilt anyway.
// Freeing a fixed variable appends a row/column to Qf and removes the implicit last T row.
// If nF would exceed the shared-memory capacity FC the vehicle is DEFERRED: its index goes to
// io.defer_list and the full-space kernel (k3_qp.cu) solves it from the untouched warm state.
//
// Shared memory (FC = 48, FB n=120 m=522): Qf 18.4 KB, packed R/T 9.2 KB, homotopy state over the
// m+n rows 21 KB, vectors 14 KB, K2's structured block 11 KB  ->  ~82 KB, 2 CTAs/SM (AB: 3).
// H*v, A*v, A'*y and single rows are matrix-free (structured block, eepacc_rows.cuh).
#include "eepacc_dev.cuh"
#include "eepacc_kernels.h"
#include "eepacc_rows.cuh"

#ifndef EEP_HOST_EMU
#ifndef K3R_NT
#define K3R_NT 128            // threads per vehicle (CTA).  Measured (profiles/README.md): 64 threads cost 3-6 % latency at equal
                              // residency and let 3 CTAs share an SM, but co-resident solves slow each other (+38 % from 1 to
                              // 3 per SM), so 2 x 128 threads and 3 x 64 threads end within 2 % (AB) and 2 x 128 wins on FB
#endif
#ifndef K3R_MINB
#define K3R_MINB 2            // CTAs per SM the register allocation is held to
#endif
#define NT K3R_NT
// The solver's primitives can be compiled as real functions (-DK3R_NOINLINE): the kernel shrinks from ~660 KB to
// ~420 KB of SASS and compiles in 1 min instead of 12, but measured 15 % slower (calls, context in local memory).
// thread-strided loops run 1-4 trips (n <= 120, m + n <= 642 over 128 threads): the default 4x unrolling only bloats
// the code (-DK3R_UNROLL1 keeps them rolled)
#ifdef K3R_UNROLL1
#define K3R_LOOP _Pragma("unroll 1")
#else
#define K3R_LOOP
#endif
#ifdef K3R_NOINLINE
#define K3R_FN __noinline__
#else
// Every primitive is inlined into the kernel: the solver context (some 60 pointers and sizes) then lives in registers.
// Left to the inliner's budget, one instantiation (ABMPC) kept a call and with it a 912-byte stack frame holding the
// context: every C.member access became a local-memory load and the kernel ran 30 % slower.
#define K3R_FN __forceinline__
#endif
#define EMU_ONLY(...)
#define LDG(p) __ldg(p)
#else
#define LDG(p) (*(p))
#define K3R_FN
#define K3R_LOOP
#endif

#define FARB      1.0e6
#define RELAX     1.0e4
#define BIGB      1.0e19
#define EPS_LI    1.0e-10
#define EPS_PD    1.0e-9
#define EPS_RATE  1.0e-12
#define EPS_ROW   1.0e-12
#define FEAS_TOL  1.0e-10
#define LP_REG    1.0e-6
#define K3R_DEFER 1000          // internal status: hand the vehicle to the full-space kernel

namespace k3r {

template <int NU, int NB>
struct CtxT {
    static constexpr int nu = NU, nb = NB;     // controls per stage, driven controls per stage
    int n, m, mt, fc, ldq;
    int nF, nX, nZ;             // free variables, fixed variables (nF + nX = n), null-space columns
    double reg, hscale;
    double tie, ynoise;         // ratio-test tie tolerance (0 for the LPs) and round-off scale of the multipliers (qp_rules.cuh)
    const double* g;                           // shared copy of the gradient [n]
    double* gw;                                // the same, writable (filled per vehicle)
    const double *lba, *uba;                   // GLOBAL: this vehicle's new row bounds [m] (far bounds for rows >= m are constants)
    const unsigned long long* rtab;            // GLOBAL (read-only path): [m] packed constraint rows (k3r_build_rowtab)
    const double *cvals, *cva, *cvb;           // [32] each: control coefficients, (a, b) of cV = a + b v_est(k)
    // factorisation
    double *Q, *M, *sub;
    // full-space vectors [n]
    double *x, *gcur, *dx, *dg, *t1, *t2, *t3, *t4;
    // fixed block, by position k in the X list [n]
    double *yX, *dyX, *xcv, *dbX;
    // reduced vectors, by Qf column / row [fc + 1]
    double *yW, *dyW, *dbw, *wv, *csA, *snA, *pF, *uF, *tF, *gF;
    // homotopy state over all rows [mt]: GLOBAL scratch of this vehicle (io.hscr), touched a few times
    // per working-set change with coalesced accesses; keeping it out of shared memory is what lets
    // four vehicles share an SM
    double *Ax, *lbc, *ubc, *Adx;
    float* anorm;
    // structured block
    double *PS, *PV, *vest, *Dc, *Cc, *Dg, *sv, *vv, *drv, *yc;
    // adjoint scratch: per-stage accumulators and per-active-row descriptors
    double *aS, *aV, *aD;       // [N+1], [N+1], [N]
    double *rS, *rV, *rv0, *rv1, *rv2;   // [fc]
    short *rk, *rj0, *rj1, *rj2, *rj3;   // [fc]
    short *flist, *fpos, *xvar, *xcon, *widx;
    signed char *st, *tabu;
    double* red;
    int* ired;
    double* tt;                 // [(2 ceil(mt/NT) + ceil(n/NT)) NT] ratios of the ratio-test candidates (pass 1 -> pass 2)
    double* hs_sm;              // shared-memory home of the homotopy state (4 mt doubles) at the first-level capacity, else null
#ifdef EEP_K3_TIMERS
    long long* ph;              // [32] clock64() cycles by phase + [32] = last stamp (thread 0 only)
#endif
};

#ifdef EEP_K3_TIMERS
#define PH(i) do { if (threadIdx.x == 0) { long long t_ = clock64(); C.ph[i] += t_ - C.ph[32]; C.ph[32] = t_; } } while (0)
#else
#define PH(i) do { } while (0)
#endif

template <class CT> __device__ __forceinline__ int mi(const CT& c, int r, int col) { return r * c.fc - (r * (r - 1)) / 2 + (col - r); }
#define QF(i, j) C.Q[(size_t)(j) * C.ldq + (i)]
#define MM(r, c) C.M[mi(C, (r), (c))]

template <class CT> __device__ __forceinline__ double lb_new(const CT& C, int c)
{
    if (c >= C.m) return -FARB;
    const double l = C.lba[c];            // plain load: K2 rewrites the bounds every step inside the fused kernel
    return (l <= -BIGB) ? -INFINITY : l;
}
template <class CT> __device__ __forceinline__ double ub_new(const CT& C, int c)
{
    if (c >= C.m) return FARB;
    const double u = C.uba[c];
    return (u >= BIGB) ? INFINITY : u;
}

// ---- block reductions (results broadcast to every thread) ----
#ifdef EEP_HOST_EMU
template <class CT> __device__ __forceinline__ double block_sum(CT&, double v) { return v; }
template <class CT> __device__ __forceinline__ double block_max(CT&, double v) { return v; }
template <class CT> __device__ __forceinline__ void block_argmin(CT&, double&, int&) {}
template <class CT> __device__ __forceinline__ void block_max3(CT&, double&, double&, double&) {}
#else
template <class CT> __device__ __forceinline__ double block_sum(CT& C, double v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) C.red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0;
    for (int w = 0; w < NT / 32; ++w) s += C.red[w];
    return s;
}
template <class CT> __device__ __forceinline__ double block_max(CT& C, double v)
{
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) C.red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = C.red[0];
    for (int w = 1; w < NT / 32; ++w) s = fmax(s, C.red[w]);
    return s;
}
// three maxima in one pass (C.red holds 16 doubles)
template <class CT> __device__ __forceinline__ void block_max3(CT& C, double& a, double& b, double& c)
{
    for (int o = 16; o > 0; o >>= 1) {
        a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o)); b = fmax(b, __shfl_xor_sync(0xffffffffu, b, o)); c = fmax(c, __shfl_xor_sync(0xffffffffu, c, o));
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { C.red[threadIdx.x >> 5] = a; C.red[NT / 32 + (threadIdx.x >> 5)] = b; C.red[2 * (NT / 32) + (threadIdx.x >> 5)] = c; }
    __syncthreads();
    a = C.red[0]; b = C.red[NT / 32]; c = C.red[2 * (NT / 32)];
    for (int w = 1; w < NT / 32; ++w) { a = fmax(a, C.red[w]); b = fmax(b, C.red[NT / 32 + w]); c = fmax(c, C.red[2 * (NT / 32) + w]); }
}
// lexicographic (t, key) minimum
template <class CT> __device__ __forceinline__ void block_argmin(CT& C, double& t, int& key)
{
    for (int o = 16; o > 0; o >>= 1) {
        double t2 = __shfl_xor_sync(0xffffffffu, t, o);
        int k2 = __shfl_xor_sync(0xffffffffu, key, o);
        if (t2 < t || (t2 == t && k2 < key)) { t = t2; key = k2; }
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { C.red[threadIdx.x >> 5] = t; C.ired[threadIdx.x >> 5] = key; }
    __syncthreads();
    t = C.red[0]; key = C.ired[0];
    for (int w = 1; w < NT / 32; ++w) {
        double t2 = C.red[w]; int k2 = C.ired[w];
        if (t2 < t || (t2 == t && k2 < key)) { t = t2; key = k2; }
    }
}
#endif

// ---- lane groups: LPC consecutive lanes of a warp share one matrix column (QR / Q accumulation) ----
#ifdef EEP_HOST_EMU
#define LPC 1
__device__ __forceinline__ double group_sum(double v) { return v; }
#else
#define LPC 4
__device__ __forceinline__ double group_sum(double v)
{
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}
#endif

struct VestSrc { const double* v_est; };

}  // namespace k3r
#include "qp_rules.cuh"
namespace k3r {

// ---- matrix-free operators: H = Psi' Hs Psi and A = Gs Psi are never formed ----
template <class CT> K3R_FN __device__ void s_prep(CT& C, const double* v)
{
    const int NH = EEP_N;
    K3R_LOOP for (int i = threadIdx.x; i < NH; i += NT) {
        double d = v[i * CT::nu];
        if (CT::nb == 2) d += v[i * CT::nu + 1];
        C.drv[i] = d;
    }
    __syncthreads();
    K3R_LOOP for (int k = threadIdx.x; k <= NH; k += NT) {
        double a = 0, b = 0;
        for (int i = 0; i < k && i < NH; ++i) { a += C.PS[k * NH + i] * C.drv[i]; b += C.PV[k * NH + i] * C.drv[i]; }
        C.sv[k] = a; C.vv[k] = b;
    }
    __syncthreads();
}
// ---- constraint rows: packed table, one 8-byte word per row c < m (built once per set-up, k3r_build_rowtab) ----
//   b0  kst (int8): stage of the state entries, -1 none, -2 jerk row (-1 on the driven controls of stage aux)
//   b1  aux: kst >= 0: bit 7 = coefficient 1 on s_k, bits 0-6 = id of the (a, b) pair, coefficient on v_k = a + b v_est(k);
//            kst == -2: the stage before
//   b2..b4  dense column of the <= 3 control entries (0xFF: none)        b5..b7  ids of their coefficients
// Far-bound rows c >= m are the unit rows e_(c-m).  No switch, no divergence: a row is one LDS.64 + table reads.
struct RowW { int kst, aux, j0, j1, j2, i0, i1, i2; };
template <class CT> __device__ __forceinline__ RowW row_load(const CT& C, int c)
{
    RowW r;
    if (c >= C.m) { r.kst = -1; r.aux = 0; r.j0 = c - C.m; r.j1 = r.j2 = 255; r.i0 = r.i1 = r.i2 = 0; return r; }
    const unsigned long long w = LDG(C.rtab + c);
    r.kst = (int)(signed char)(w & 0xff); r.aux = (int)((w >> 8) & 0xff);
    r.j0 = (int)((w >> 16) & 0xff); r.j1 = (int)((w >> 24) & 0xff); r.j2 = (int)((w >> 32) & 0xff);
    r.i0 = (int)((w >> 40) & 0xff); r.i1 = (int)((w >> 48) & 0xff); r.i2 = (int)((w >> 56) & 0xff);
    return r;
}
template <class CT> __device__ __forceinline__ double row_cV(const CT& C, const RowW& r)
{
    const int id = r.aux & 127;
    return C.cva[id] + C.cvb[id] * C.vest[r.kst];
}
// one entry A(c, j)
template <class CT> __device__ __forceinline__ double s_entry(const CT& C, const RowW& r, int j)
{
    const int NH = EEP_N;
    const int i = j / CT::nu, cj = j % CT::nu;
    double val = 0.0;
    if (r.kst >= 0 && cj < CT::nb && i < r.kst) {
        val = row_cV(C, r) * C.PV[r.kst * NH + i];
        if (r.aux & 128) val += C.PS[r.kst * NH + i];
    }
    if (r.j0 == j) val += C.cvals[r.i0];
    if (r.j1 == j) val += C.cvals[r.i1];
    if (r.j2 == j) val += C.cvals[r.i2];
    if (r.kst == -2 && cj < CT::nb && i == r.aux) val -= 1.0;
    return val;
}
// a_c' v for a vector whose state sensitivities are in sv/vv/drv (s_prep)
template <class CT> __device__ __forceinline__ double s_rowdot(const CT& C, const RowW& r, const double* v)
{
    double val = 0.0;
    if (r.kst >= 0) {
        val = row_cV(C, r) * C.vv[r.kst];
        if (r.aux & 128) val += C.sv[r.kst];
    }
    if (r.j0 != 255) val += C.cvals[r.i0] * v[r.j0];
    if (r.j1 != 255) val += C.cvals[r.i1] * v[r.j1];
    if (r.j2 != 255) val += C.cvals[r.i2] * v[r.j2];
    if (r.kst == -2) val -= C.drv[r.aux];
    return val;
}
// unit row?  -> variable index and coefficient
template <class CT> __device__ __forceinline__ bool s_unit(const CT& C, const RowW& r, int& jv, double& cv)
{
    if (r.kst == -1 && r.j1 == 255 && r.j0 != 255) { jv = r.j0; cv = C.cvals[r.i0]; return true; }
    return false;
}

// full row (all n entries)
template <class CT> K3R_FN __device__ void get_row(CT& C, int c, double* a)
{
    const RowW rd = row_load(C, c);
    K3R_LOOP for (int j = threadIdx.x; j < C.n; j += NT) a[j] = s_entry(C, rd, j);
    __syncthreads();
}
template <class CT> K3R_FN __device__ void Hmul(CT& C, const double* v, double* out)
{
    const int n = C.n;
    const int NH = EEP_N, nu = CT::nu, nb = CT::nb;
    s_prep(C, v);
    // y_k on the core [v, driven]: D_k w_k + C_k w_{k-1} + C_{k+1}' w_{k+1}
    K3R_LOOP for (int k = threadIdx.x; k < NH; k += NT) {
        double w0[3] = { C.vv[k], v[k * nu], nb == 2 ? v[k * nu + 1] : 0.0 };
        double y[3] = { 0, 0, 0 };
        const double* D = C.Dc + k * 9;
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) y[a] += D[a * 3 + b] * w0[b];
        if (k > 0) {
            double w1[3] = { C.vv[k - 1], v[(k - 1) * nu], nb == 2 ? v[(k - 1) * nu + 1] : 0.0 };
            const double* Cm = C.Cc + k * 9;
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) y[a] += Cm[a * 3 + b] * w1[b];
        }
        if (k + 1 < NH) {
            double w2[3] = { C.vv[k + 1], v[(k + 1) * nu], nb == 2 ? v[(k + 1) * nu + 1] : 0.0 };
            const double* Cp = C.Cc + (k + 1) * 9;
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) y[a] += Cp[b * 3 + a] * w2[b];
        }
        C.yc[k * 3] = y[0]; C.yc[k * 3 + 1] = y[1]; C.yc[k * 3 + 2] = y[2];
    }
    __syncthreads();
    K3R_LOOP for (int j = threadIdx.x; j < n; j += NT) {
        const int i = j / nu, cj = j % nu;
        double val;
        if (cj < nb) {
            val = C.yc[i * 3 + 1 + cj];
            double acc = 0.0;
            for (int k = i + 1; k < NH; ++k) acc += C.PV[k * NH + i] * C.yc[k * 3];
            val += acc;
        } else val = C.Dg[i * 4 + (cj - nb)] * v[j];
        out[j] = val + C.reg * v[j];
    }
    __syncthreads();
}
template <class CT> K3R_FN __device__ void Amul(CT& C, const double* v, double* out)
{
    const int n = C.n, m = C.m;
    s_prep(C, v);
    K3R_LOOP for (int r = threadIdx.x; r < m; r += NT) {
        out[r] = s_rowdot(C, row_load(C, r), v);
    }
    K3R_LOOP for (int j = threadIdx.x; j < n; j += NT) out[m + j] = v[j];
    __syncthreads();
}
// w[j] = sum over the active general rows r in [nZ, nF) of A(widx[r], j) yv[r]
//        (+ cv_k yXv[k] at j = xvar[k] when yXv != nullptr).  Deterministic gather, no atomics.
template <class CT> K3R_FN __device__ void ATmul(CT& C, const double* yv, const double* yXv, double* w)
{
    const int n = C.n, NH = EEP_N, nu = CT::nu, nb = CT::nb, nZ = C.nZ, nF = C.nF;
    __syncthreads();
    K3R_LOOP for (int r = nZ + threadIdx.x; r < nF; r += NT) {
        const RowW rd = row_load(C, C.widx[r]);
        const double y = yv[r];
        C.rk[r] = (short)rd.kst;
        C.rS[r] = (rd.kst == -2 || (rd.aux & 128)) ? y : 0.0;     // jerk rows: -y on the driven controls of the stage before
        C.rV[r] = (rd.kst >= 0) ? row_cV(C, rd) * y : 0.0;
        C.rj0[r] = (short)(rd.kst == -2 ? 256 + rd.aux : rd.j0);  // jerk rows carry the stage before here (their own entries follow)
        C.rj1[r] = (short)rd.j1; C.rj2[r] = (short)rd.j2;
        C.rv0[r] = (rd.j0 != 255) ? C.cvals[rd.i0] * y : 0.0;
        C.rv1[r] = (rd.j1 != 255) ? C.cvals[rd.i1] * y : 0.0;
        C.rv2[r] = (rd.j2 != 255) ? C.cvals[rd.i2] * y : 0.0;
        C.rj3[r] = (short)rd.j0;
    }
    __syncthreads();
    // per-stage accumulators: thread k gathers its stage (oldest row first)
    K3R_LOOP for (int k = threadIdx.x; k <= NH; k += NT) {
        double s = 0, v = 0, d = 0;
        for (int r = nF - 1; r >= nZ; --r) {
            const int kr = C.rk[r];
            if (kr == k) { s += C.rS[r]; v += C.rV[r]; }
            else if (kr == -2 && C.rj0[r] - 256 == k) d += C.rS[r];
        }
        C.aS[k] = s; C.aV[k] = v;
        if (k < NH) C.aD[k] = d;
    }
    __syncthreads();
    K3R_LOOP for (int j = threadIdx.x; j < n; j += NT) {
        const int i = j / nu, cj = j % nu;
        double val = 0.0;
        for (int r = nF - 1; r >= nZ; --r) {           // oldest first
            if (C.rj3[r] == j) val += C.rv0[r];
            if (C.rj1[r] == j) val += C.rv1[r];
            if (C.rj2[r] == j) val += C.rv2[r];
        }
        if (cj < nb) {
            double acc = 0.0;
            for (int k = i + 1; k <= NH; ++k) acc += C.aS[k] * C.PS[k * NH + i] + C.aV[k] * C.PV[k * NH + i];
            val += acc - C.aD[i];
        }
        w[j] = val;
    }
    __syncthreads();
    if (yXv) {
        K3R_LOOP for (int k = threadIdx.x; k < C.nX; k += NT) w[C.xvar[k]] += C.xcv[k] * yXv[k];
        __syncthreads();
    }
}

// ---- reduced dense primitives; every routine ends with data visible after a __syncthreads() ----
template <class CT> K3R_FN __device__ void gatherF(CT& C, const double* a)
{
    K3R_LOOP for (int i = threadIdx.x; i < C.nF; i += NT) C.gF[i] = a[C.flist[i]];
    __syncthreads();
}
// w[j] = qf_j' aF for j in [c0, c1), aF compact [nF]
template <class CT> K3R_FN __device__ void Qtmul(CT& C, const double* aF, double* w, int c0, int c1)
{
    const int nF = C.nF;
    K3R_LOOP for (int j = c0 + threadIdx.x; j < c1; j += NT) {
        double s0 = 0, s1 = 0;
        int i = 0;
        for (; i + 1 < nF; i += 2) { s0 += QF(i, j) * aF[i]; s1 += QF(i + 1, j) * aF[i + 1]; }
        if (i < nF) s0 += QF(i, j) * aF[i];
        w[j] = s0 + s1;
    }
    __syncthreads();
}
// outF[i] = sum_{c in [c0,c1)} Qf(i,c) p[c], i < nF (compact)
template <class CT> K3R_FN __device__ void Qmul(CT& C, const double* p, int c0, int c1, double* outF)
{
    const int nF = C.nF;
    K3R_LOOP for (int i = threadIdx.x; i < nF; i += NT) {
        double s0 = 0, s1 = 0;
        int c = c0;
        for (; c + 1 < c1; c += 2) { s0 += QF(i, c) * p[c]; s1 += QF(i, c + 1) * p[c + 1]; }
        if (c < c1) s0 += QF(i, c) * p[c];
        outF[i] = s0 + s1;
    }
    __syncthreads();
}
// solve M[lo:hi,lo:hi] x = b (back substitution); b is destroyed, x may not alias b
template <class CT> K3R_FN __device__ void usolve(CT& C, int lo, int hi, double* b, double* x)
{
#ifdef EEP_HOST_EMU
    for (int r = hi - 1; r >= lo; --r) {
        double s = b[r];
        for (int c = r + 1; c < hi; ++c) s -= MM(r, c) * x[c];
        x[r] = s * (1.0 / MM(r, r));
    }
#else
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int b1 = hi; b1 > lo; b1 -= 32) {
        const int b0 = (b1 - 32 > lo) ? b1 - 32 : lo, nbk = b1 - b0;
        __syncthreads();
        if (warp == 0) {
            const int r = b0 + lane;
            double acc = (lane < nbk) ? b[r] : 0.0;
            const double inv = (lane < nbk) ? 1.0 / MM(r, r) : 0.0;
            for (int t = nbk - 1; t >= 0; --t) {
                const double xt = __shfl_sync(0xffffffffu, acc * inv, t);
                if (lane < t) acc -= MM(r, b0 + t) * xt;
                if (lane == t) x[r] = xt;
            }
        }
        __syncthreads();
        K3R_LOOP for (int i = lo + threadIdx.x; i < b0; i += NT) {
            double s0 = 0, s1 = 0;
            int c = b0;
            for (; c + 1 < b1; c += 2) { s0 += MM(i, c) * x[c]; s1 += MM(i, c + 1) * x[c + 1]; }
            if (c < b1) s0 += MM(i, c) * x[c];
            b[i] -= s0 + s1;
        }
    }
    __syncthreads();
#endif
}
// solve M[lo:hi,lo:hi]' y = v (forward substitution); v destroyed
template <class CT> K3R_FN __device__ void tsolve(CT& C, int lo, int hi, double* v, double* y)
{
#ifdef EEP_HOST_EMU
    for (int r = lo; r < hi; ++r) {
        double s = v[r];
        for (int c = lo; c < r; ++c) s -= MM(c, r) * y[c];
        y[r] = s * (1.0 / MM(r, r));
    }
#else
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int b0 = lo; b0 < hi; b0 += 32) {
        const int b1 = (b0 + 32 < hi) ? b0 + 32 : hi, nbk = b1 - b0;
        __syncthreads();
        if (warp == 0) {
            const int r = b0 + lane;
            double acc = (lane < nbk) ? v[r] : 0.0;
            const double inv = (lane < nbk) ? 1.0 / MM(r, r) : 0.0;
            for (int t = 0; t < nbk; ++t) {
                const double yt = __shfl_sync(0xffffffffu, acc * inv, t);
                if (lane > t && lane < nbk) acc -= MM(b0 + t, r) * yt;
                if (lane == t) y[r] = yt;
            }
        }
        __syncthreads();
        K3R_LOOP for (int i = b1 + threadIdx.x; i < hi; i += NT) {
            double s0 = 0, s1 = 0;
            int c = b0;
            for (; c + 1 < b1; c += 2) { s0 += MM(c, i) * y[c]; s1 += MM(c + 1, i) * y[c + 1]; }
            if (c < b1) s0 += MM(c, i) * y[c];
            v[i] -= s0 + s1;
        }
    }
    __syncthreads();
#endif
}

// Givens sweep that adds the row whose free part is aF (compact).  nZt = null-space columns that take
// part (0..nZt-1), nR = dimension of the valid R block to keep updated (0: none).  The new
// range-space column / T row is nZt-1.  Returns 1 if the row is linearly dependent (no change).
template <class CT> K3R_FN __device__ int givens_add(CT& C, int cidx, int side, const double* aF, double an2, int nZt, int nR, bool check_li)
{
    const int nF = C.nF;
    Qtmul(C, aF, C.wv, 0, nF);
    if (check_li) {
        double zn = 0;
        K3R_LOOP for (int j = threadIdx.x; j < nZt; j += NT) zn += C.wv[j] * C.wv[j];
        zn = block_sum(C, zn);
        const double san = sqrt(an2);
        if (nZt == 0 || sqrt(zn) <= EPS_LI * (san > 1.0 ? san : 1.0)) return 1;
    }
    if (threadIdx.x == 0) {
        double carry = C.wv[0];
        for (int k = 0; k < nZt - 1; ++k) {
            double p = carry, q = C.wv[k + 1];
            double r = hypot(p, q);
            double cs = 1.0, sn = 0.0;
            if (r != 0.0) { cs = q / r; sn = p / r; carry = r; } else carry = q;
            C.csA[k] = cs; C.snA[k] = sn;
            C.wv[k] = 0.0;
        }
        if (nZt > 0) C.wv[nZt - 1] = carry;
    }
    __syncthreads();
    // Qf: one row per thread, rotations applied left to right with the carry in a register
    K3R_LOOP for (int i = threadIdx.x; i < nF; i += NT) {
        double u = QF(i, 0);
        for (int k = 0; k < nZt - 1; ++k) {
            double v = QF(i, k + 1);
            double cs = C.csA[k], sn = C.snA[k];
            QF(i, k) = cs * u - sn * v;
            u = sn * u + cs * v;
        }
        if (nZt > 0) QF(i, nZt - 1) = u;
    }
    if (nR > 1) {
        // R phase A: column rotations k = 0..nR-2, one row of R per thread
        K3R_LOOP for (int i = threadIdx.x; i < nR; i += NT) {
            int k0 = i > 0 ? i - 1 : 0;
            double u;
            if (i > 0) {        // rotation k = i-1 mixes (i,i-1)=0 with (i,i)
                double d = MM(i, i);
                C.sub[i] = -C.snA[i - 1] * d;
                u = C.csA[i - 1] * d;
                k0 = i;
            } else u = MM(0, 0);
            for (int k = k0; k < nR - 1; ++k) {
                double v = MM(i, k + 1);
                double cs = C.csA[k], sn = C.snA[k];
                MM(i, k) = cs * u - sn * v;
                u = sn * u + cs * v;
            }
            MM(i, nR - 1) = u;
        }
        // R phase B: re-triangularise (row rotations k = 0..nR-2 on rows k, k+1).  Column j only sees rotations
        // 0..j, one after the other, and rotation k is defined by column k once rotations 0..k-1 have passed over it:
        // ONE WARP walks the chain with lane j owning column j (shuffles, no block barrier per rotation).
#ifndef EEP_HOST_EMU
        __syncthreads();
        if (threadIdx.x < 32) warp_chain_rows(C, nR);
#else
        for (int k = 0; k < nR - 1; ++k) {
            __syncthreads();
            double pa = MM(k, k), pb = C.sub[k + 1];
            double rr = hypot(pa, pb);
            __syncthreads();
            if (rr != 0.0) {
                double c2 = pa / rr, s2 = pb / rr;
                K3R_LOOP for (int j = k + threadIdx.x; j < nR; j += NT) {
                    double u = MM(k, j);
                    double v = (j == k) ? pb : MM(k + 1, j);
                    MM(k, j) = c2 * u + s2 * v;
                    if (j > k) MM(k + 1, j) = -s2 * u + c2 * v;
                }
            }
        }
#endif
    }
    __syncthreads();
    // new T row = M row nZt-1 (columns >= nZt-1)
    const int rn = nZt - 1;
    K3R_LOOP for (int c = rn + threadIdx.x; c < nF; c += NT) MM(rn, c) = C.wv[c];
    if (threadIdx.x == 0) { C.widx[rn] = (short)cidx; C.st[cidx] = (signed char)side; }
    C.nZ = rn;
    __syncthreads();
    return 0;
}

// Remove the general row that owns Qf column rp.  Afterwards column nZ_old is the new null-space
// direction; R is not touched.  yW / widx are shifted with the rows.
template <class CT> K3R_FN __device__ void remove_general(CT& C, int rp)
{
    const int nF = C.nF, nZ = C.nZ;
    if (threadIdx.x == 0) C.st[C.widx[rp]] = 0;
    __syncthreads();
    // shift rows nZ..rp-1 one row down the diagonal (row r-1 -> row r); one column per thread
    K3R_LOOP for (int c = nZ + threadIdx.x; c < nF; c += NT) {
        if (c < rp) C.sub[c + 1] = MM(c, c);
        int rtop = c < rp ? c : rp;
        for (int r = rtop; r > nZ; --r) MM(r, c) = MM(r - 1, c);
    }
    if (threadIdx.x == 0) {
        for (int r = rp; r > nZ; --r) { C.widx[r] = C.widx[r - 1]; C.yW[r] = C.yW[r - 1]; }
    }
    // Givens wavefront: rotation at r (columns r-1, r) for r = rp .. nZ+1.  Row rr only sees rotations rp..rr, one
    // after the other, and rotation r is defined by row r once rotations rp..r+1 have passed over it: ONE WARP
    // walks the chain with a lane per row (shuffles instead of two block barriers per rotation).
#ifndef EEP_HOST_EMU
    __syncthreads();
    if (threadIdx.x < 32) warp_chain_cols(C, rp, nZ);
#else
    for (int r = rp; r > nZ; --r) {
        __syncthreads();
        double a = C.sub[r], b = MM(r, r);
        double rho = hypot(a, b);
        double cs = 1.0, sn = 0.0;
        if (rho != 0.0) { cs = b / rho; sn = a / rho; }
        __syncthreads();
        if (threadIdx.x == 0) { C.csA[r] = cs; C.snA[r] = sn; }
        // rows r' in [nZ+1, r]: elements (r', r-1) and (r', r)
        K3R_LOOP for (int rr = nZ + 1 + threadIdx.x; rr <= r; rr += NT) {
            double u = (rr == r) ? a : MM(rr, r - 1);
            double v = MM(rr, r);
            if (rr < r) MM(rr, r - 1) = cs * u - sn * v;
            MM(rr, r) = sn * u + cs * v;
        }
    }
#endif
    __syncthreads();
    K3R_LOOP for (int i = threadIdx.x; i < nF; i += NT) {
        if (rp > nZ) {
            double v = QF(i, rp);
            for (int r = rp; r > nZ; --r) {
                double u = QF(i, r - 1);
                double cs = C.csA[r], sn = C.snA[r];
                QF(i, r) = sn * u + cs * v;
                v = cs * u - sn * v;          // becomes column r-1, the right-hand column of the next rotation
            }
            QF(i, nZ) = v;
        }
    }
    C.nZ = nZ + 1;
    __syncthreads();
}

// Free the variable pinned by X entry k (remove its unit row from the working set): append the
// coordinate to Qf and T (implicit last row of the full T made explicit), then remove that row.
// Returns 1 if the capacity is exhausted (nothing changed).
template <class CT> K3R_FN __device__ int free_var(CT& C, int k)
{
    const int nF = C.nF, nZ = C.nZ;
    if (nF >= C.fc) return 1;
    __syncthreads();
    const int j = C.xvar[k], c = C.xcon[k];
    const double cv = C.xcv[k], yk = C.yX[k];
    __syncthreads();
    K3R_LOOP for (int i = threadIdx.x; i <= nF; i += NT) { QF(i, nF) = (i == nF) ? 1.0 : 0.0; QF(nF, i) = (i == nF) ? 1.0 : 0.0; }
    K3R_LOOP for (int r = nZ + threadIdx.x; r < nF; r += NT) {
        const RowW rd = row_load(C, C.widx[r]);
        MM(r, nF) = s_entry(C, rd, j);
    }
    if (threadIdx.x == 0) {
        MM(nF, nF) = cv;
        C.widx[nF] = (short)c; C.yW[nF] = yk;
        C.flist[nF] = (short)j; C.fpos[j] = (short)nF;
        for (int q = k; q + 1 < C.nX; ++q) {
            C.xvar[q] = C.xvar[q + 1]; C.xcon[q] = C.xcon[q + 1]; C.xcv[q] = C.xcv[q + 1]; C.yX[q] = C.yX[q + 1];
            C.fpos[C.xvar[q]] = (short)(-1 - q);
        }
    }
    C.nF = nF + 1; C.nX -= 1;
    __syncthreads();
    remove_general(C, nF);
    return 0;
}

// OUT_g = H V_g for g < G vectors at once (same arithmetic as Hmul per vector; the stage phases of the G
// vectors share the barriers).  V, OUT: [G][n] (OUT may alias V: the last phase reads V only at the entry
// it writes); scr: [G][HG_SCR] doubles.
#define HG_SCR (4 * EEP_N + 2 * (EEP_N + 1))
template <class CT> K3R_FN __device__ void HmulG(CT& C, int G, const double* V, double* OUT, double* scr)
{
    const int n = C.n, NH = EEP_N, nu = CT::nu, nb = CT::nb;
    K3R_LOOP for (int e = threadIdx.x; e < G * NH; e += NT) {
        const int g = e / NH, i = e - g * NH;
        const double* v = V + g * n;
        double d = v[i * nu];
        if (nb == 2) d += v[i * nu + 1];
        scr[g * HG_SCR + i] = d;                                    // drv
    }
    __syncthreads();
    K3R_LOOP for (int e = threadIdx.x; e < G * (NH + 1); e += NT) {
        const int g = e / (NH + 1), k = e - g * (NH + 1);
        const double* drv = scr + g * HG_SCR;
        double a = 0, b = 0;
        for (int i = 0; i < k && i < NH; ++i) { a += C.PS[k * NH + i] * drv[i]; b += C.PV[k * NH + i] * drv[i]; }
        scr[g * HG_SCR + NH + k] = b;                               // vv (sv does not enter H v)
    }
    __syncthreads();
    K3R_LOOP for (int e = threadIdx.x; e < G * NH; e += NT) {
        const int g = e / NH, k = e - g * NH;
        const double* v = V + g * n;
        const double* vv = scr + g * HG_SCR + NH;
        double* yc = scr + g * HG_SCR + 2 * NH + 1;
        double w0[3] = { vv[k], v[k * nu], nb == 2 ? v[k * nu + 1] : 0.0 };
        double y[3] = { 0, 0, 0 };
        const double* D = C.Dc + k * 9;
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) y[a] += D[a * 3 + b] * w0[b];
        if (k > 0) {
            double w1[3] = { vv[k - 1], v[(k - 1) * nu], nb == 2 ? v[(k - 1) * nu + 1] : 0.0 };
            const double* Cm = C.Cc + k * 9;
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) y[a] += Cm[a * 3 + b] * w1[b];
        }
        if (k + 1 < NH) {
            double w2[3] = { vv[k + 1], v[(k + 1) * nu], nb == 2 ? v[(k + 1) * nu + 1] : 0.0 };
            const double* Cp = C.Cc + (k + 1) * 9;
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) y[a] += Cp[b * 3 + a] * w2[b];
        }
        yc[k * 3] = y[0]; yc[k * 3 + 1] = y[1]; yc[k * 3 + 2] = y[2];
    }
    __syncthreads();
    K3R_LOOP for (int e = threadIdx.x; e < G * n; e += NT) {
        const int g = e / n, j = e - g * n;
        const double* v = V + g * n;
        const double* yc = scr + g * HG_SCR + 2 * NH + 1;
        const int i = j / nu, cj = j % nu;
        double val;
        if (cj < nb) {
            val = yc[i * 3 + 1 + cj];
            double acc = 0.0;
            for (int k = i + 1; k < NH; ++k) acc += C.PV[k * NH + i] * yc[k * 3];
            val += acc;
        } else val = C.Dg[i * 4 + (cj - nb)] * v[j];
        OUT[g * n + j] = val + C.reg * v[j];
    }
    __syncthreads();
}

// R = chol(Z' H Z) from scratch on the first nZ columns; returns 1 if not positive definite.
// H Z is formed up to four null-space columns at a time, in place, in the nine work vectors gcur .. dbX
// (9 n contiguous doubles; none of them is live between warm_setup and setup_aux).
template <class CT> K3R_FN __device__ int refactor_R(CT& C)
{
    const int n = C.n, nF = C.nF, nZ = C.nZ;
    int GMAX = (9 * n) / (n + HG_SCR);
    if (GMAX > 4) GMAX = 4;
    double* V = C.gcur;
#ifndef EEP_HOST_EMU
    if (C.hs_sm) {
        // first-level capacity: the ratio-test slots and the homotopy state behind them are dead until setup_aux and hold
        // up to 16 null-space columns at a time (fewer barrier-separated passes of the stage operators)
        const int navail = (2 * ((C.mt + NT - 1) / NT) + (n + NT - 1) / NT) * NT + 4 * C.mt;
        GMAX = navail / (n + HG_SCR);
        if (GMAX > 16) GMAX = 16;
        V = C.tt;
    }
#endif
    double* scr = V + GMAX * n;
    for (int j0 = 0; j0 < nZ; j0 += GMAX) {
        const int G = (nZ - j0 < GMAX) ? nZ - j0 : GMAX;
        K3R_LOOP for (int e = threadIdx.x; e < G * n; e += NT) V[e] = 0.0;
        __syncthreads();
        K3R_LOOP for (int e = threadIdx.x; e < G * nF; e += NT) {
            const int g = e / nF, i = e - g * nF;
            V[g * n + C.flist[i]] = QF(i, j0 + g);
        }
        __syncthreads();
        HmulG(C, G, V, V, scr);
        // M(i, j0 + g) = z_i' (H z_(j0+g)),  i <= j0 + g
        const int rows = j0 + G;
        K3R_LOOP for (int e = threadIdx.x; e < G * rows; e += NT) {
            const int g = e / rows, i = e - g * rows;
            if (i > j0 + g) continue;
            const double* hz = V + g * n;
            double s0 = 0, s1 = 0;
            int f = 0;
            for (; f + 1 < nF; f += 2) { s0 += QF(f, i) * hz[C.flist[f]]; s1 += QF(f + 1, i) * hz[C.flist[f + 1]]; }
            if (f < nF) s0 += QF(f, i) * hz[C.flist[f]];
            MM(i, j0 + g) = s0 + s1;
        }
        __syncthreads();
    }
    const double thr = EPS_PD * C.hscale;
#ifndef EEP_HOST_EMU
    if (nZ <= 32) {
        // one warp, lane j owns column j: left-looking upper Cholesky, row by row; R(i,i) is handed on by shuffle
        __syncthreads();
        int bad = 0;
        if (threadIdx.x < 32) {
            const int j = threadIdx.x;
            for (int i = 0; i < nZ; ++i) {
                double v = 0.0;
                if (j >= i && j < nZ) {
                    v = MM(i, j);
                    for (int k = 0; k < i; ++k) v -= MM(k, i) * MM(k, j);
                }
                const double d = __shfl_sync(0xffffffffu, v, i);
                if (!(d > thr)) { bad = 1; break; }
                const double rk = sqrt(d), inv = 1.0 / rk;
                if (j == i) MM(i, i) = rk;
                else if (j > i && j < nZ) MM(i, j) = v * inv;
                __syncwarp();
            }
            if (j == 0) C.ired[7] = bad;
        }
        __syncthreads();
        return C.ired[7];
    }
#endif
    for (int k = 0; k < nZ; ++k) {
        __syncthreads();
        double d = MM(k, k);
        if (!(d > thr)) return 1;
        double rk = sqrt(d), inv = 1.0 / rk;
        const int w = nZ - 1 - k;                 // trailing dimension
        K3R_LOOP for (int e = threadIdx.x; e < w * w; e += NT) {
            int i = k + 1 + e / w, j = k + 1 + e % w;
            if (j >= i) MM(i, j) -= (MM(k, i) * inv) * (MM(k, j) * inv);
        }
        __syncthreads();
        K3R_LOOP for (int j = k + threadIdx.x; j < nZ; j += NT) MM(k, j) = (j == k) ? rk : MM(k, j) * inv;
    }
    __syncthreads();
    return 0;
}

// KKT step: H dx + dg = A_W' dy,  A_W dx = db.  db / dy: dbw,dyW by Qf column (general rows) and
// dbX,dyX by X position (unit rows of the fixed variables).
template <class CT> K3R_FN __device__ void kkt_step(CT& C, const double* dg, const double* dbw, const double* dbX, double* dx, double* dyW, double* dyX)
{
    const int n = C.n, nF = C.nF, nZ = C.nZ, nX = C.nX, nAC = nF - nZ;
    double* u = C.t2;
    K3R_LOOP for (int i = threadIdx.x; i < n; i += NT) dx[i] = 0.0;
    __syncthreads();
    K3R_LOOP for (int k = threadIdx.x; k < nX; k += NT) dx[C.xvar[k]] = dbX[k] / C.xcv[k];
    __syncthreads();
    if (nAC > 0) {
        s_prep(C, dx);
        K3R_LOOP for (int r = nZ + threadIdx.x; r < nF; r += NT) {
            const RowW rd = row_load(C, C.widx[r]);
            C.tF[r] = dbw[r] - s_rowdot(C, rd, dx);
        }
        __syncthreads();
        usolve(C, nZ, nF, C.tF, C.pF);
        Qmul(C, C.pF, nZ, nF, C.uF);
        K3R_LOOP for (int i = threadIdx.x; i < nF; i += NT) dx[C.flist[i]] = C.uF[i];
        __syncthreads();
    }
    PH(7);
    Hmul(C, dx, u);
    K3R_LOOP for (int i = threadIdx.x; i < n; i += NT) u[i] += dg[i];
    __syncthreads();
    PH(8);
    if (nZ > 0) {
        gatherF(C, u);
        Qtmul(C, C.gF, C.tF, 0, nZ);
        K3R_LOOP for (int j = threadIdx.x; j < nZ; j += NT) C.tF[j] = -C.tF[j];
        __syncthreads();
        tsolve(C, 0, nZ, C.tF, C.pF);
        usolve(C, 0, nZ, C.pF, C.tF);
        Qmul(C, C.tF, 0, nZ, C.uF);              // p = Z q (compact)
        K3R_LOOP for (int i = threadIdx.x; i < n; i += NT) C.t1[i] = 0.0;
        __syncthreads();
        K3R_LOOP for (int i = threadIdx.x; i < nF; i += NT) { const int j = C.flist[i]; C.t1[j] = C.uF[i]; dx[j] += C.uF[i]; }
        __syncthreads();
        PH(9);
        Hmul(C, C.t1, C.t3);
        K3R_LOOP for (int i = threadIdx.x; i < n; i += NT) u[i] += C.t3[i];
        __syncthreads();
        PH(10);
    }
    if (nAC > 0) {
        gatherF(C, u);
        Qtmul(C, C.gF, C.tF, nZ, nF);
        tsolve(C, nZ, nF, C.tF, dyW);
    }
    PH(11);
    if (nX > 0) {
        ATmul(C, dyW, nullptr, C.t3);
        K3R_LOOP for (int k = threadIdx.x; k < nX; k += NT) { const int j = C.xvar[k]; dyX[k] = (u[j] - C.t3[j]) / C.xcv[k]; }
        __syncthreads();
    }
    PH(12);
}

// g_cur = A_W' y - H x  (optionally recomputing y by least squares first), duals sign-clamped
template <class CT> K3R_FN __device__ void reanchor(CT& C, bool recompute_y)
{
    const int n = C.n, nF = C.nF, nZ = C.nZ, nX = C.nX;
    Hmul(C, C.x, C.t1);
    if (recompute_y && (nX + nF - nZ) > 0) {
        K3R_LOOP for (int i = threadIdx.x; i < n; i += NT) C.t2[i] = C.t1[i] + C.gcur[i];
        __syncthreads();
        if (nF > nZ) {
            gatherF(C, C.t2);
            Qtmul(C, C.gF, C.tF, nZ, nF);
            tsolve(C, nZ, nF, C.tF, C.yW);
        }
        if (nX > 0) {
            ATmul(C, C.yW, nullptr, C.t3);
            K3R_LOOP for (int k = threadIdx.x; k < nX; k += NT) { const int j = C.xvar[k]; C.yX[k] = (C.t2[j] - C.t3[j]) / C.xcv[k]; }
            __syncthreads();
        }
        double ym = C.ynoise;
        K3R_LOOP for (int r = nZ + threadIdx.x; r < nF; r += NT) ym = fmax(ym, fabs(C.yW[r]));
        K3R_LOOP for (int k = threadIdx.x; k < nX; k += NT) ym = fmax(ym, fabs(C.yX[k]));
        C.ynoise = block_max(C, ym);
    }
    K3R_LOOP for (int r = nZ + threadIdx.x; r < nF; r += NT) {
        int c = C.widx[r];
        if (lb_new(C, c) != ub_new(C, c)) {
            if (C.st[c] < 0 && C.yW[r] < 0) C.yW[r] = 0;
            if (C.st[c] > 0 && C.yW[r] > 0) C.yW[r] = 0;
        }
    }
    K3R_LOOP for (int k = threadIdx.x; k < nX; k += NT) {
        int c = C.xcon[k];
        if (lb_new(C, c) != ub_new(C, c)) {
            if (C.st[c] < 0 && C.yX[k] < 0) C.yX[k] = 0;
            if (C.st[c] > 0 && C.yX[k] > 0) C.yX[k] = 0;
        }
    }
    __syncthreads();
    ATmul(C, C.yW, C.yX, C.t3);
    K3R_LOOP for (int j = threadIdx.x; j < n; j += NT) C.gcur[j] = C.t3[j] - C.t1[j];
    __syncthreads();
}

// compact free part of row c into aF (and its full squared norm)
template <class CT> K3R_FN __device__ double get_rowF(CT& C, int c, double* aF)
{
    const RowW rd = row_load(C, c);
    double s2 = 0;
    K3R_LOOP for (int j = threadIdx.x; j < C.n; j += NT) {
        const double v = s_entry(C, rd, j);
        s2 += v * v;
        const int f = C.fpos[j];
        if (f >= 0) aF[f] = v;
    }
    s2 = block_sum(C, s2);
    __syncthreads();
    return s2;
}

// see oracle/qp_oracle.c:resolve_pending
template <class CT> K3R_FN __device__ int resolve_pending(CT& C, int c_out, int side_out, unsigned long long& n_exch, bool zero_step)
{
    const int n = C.n, mt = C.mt;
    bool first = true, moved = false;
    for (;;) {
        const int kpd = C.nZ - 1;
        if (kpd < 0) break;
        const int nF = C.nF;
        double *zf = C.t4, *hz = C.t1, *d = C.dx;
        K3R_LOOP for (int i = threadIdx.x; i < n; i += NT) zf[i] = 0.0;
        __syncthreads();
        K3R_LOOP for (int i = threadIdx.x; i < nF; i += NT) zf[C.flist[i]] = QF(i, kpd);
        __syncthreads();
        Hmul(C, zf, hz);
        double zhz = 0;
        K3R_LOOP for (int i = threadIdx.x; i < n; i += NT) zhz += zf[i] * hz[i];
        zhz = block_sum(C, zhz);
        gatherF(C, hz);
        Qtmul(C, C.gF, C.tF, 0, kpd);
        tsolve(C, 0, kpd, C.tF, C.pF);           // r = R'^-1 Z'Hz
        double rr = 0;
        K3R_LOOP for (int j = threadIdx.x; j < kpd; j += NT) rr += C.pF[j] * C.pF[j];
        rr = block_sum(C, rr);
        double rho2 = zhz - rr;
        if (rho2 > EPS_PD * C.hscale) {
            K3R_LOOP for (int j = threadIdx.x; j < kpd; j += NT) MM(j, kpd) = C.pF[j];
            if (threadIdx.x == 0) MM(kpd, kpd) = sqrt(rho2);
            __syncthreads();
            break;
        }
        K3R_LOOP for (int j = threadIdx.x; j < kpd; j += NT) C.tF[j] = C.pF[j];
        __syncthreads();
        usolve(C, 0, kpd, C.tF, C.uF);           // q
        Qmul(C, C.uF, 0, kpd, C.pF);             // Z q (compact)
        K3R_LOOP for (int i = threadIdx.x; i < n; i += NT) d[i] = 0.0;
        __syncthreads();
        K3R_LOOP for (int i = threadIdx.x; i < nF; i += NT) { const int j = C.flist[i]; d[j] = zf[j] - C.pF[i]; }
        __syncthreads();
        double sgn = 0;
        if (first) {
            get_row(C, c_out, C.t2);
            K3R_LOOP for (int i = threadIdx.x; i < n; i += NT) sgn += C.t2[i] * d[i];
            sgn = block_sum(C, sgn);
            if (side_out > 0) sgn = -sgn;
        } else {
            Hmul(C, C.x, hz);
            double gn = 0, dn2 = 0;
            K3R_LOOP for (int i = threadIdx.x; i < n; i += NT) { const double gi = hz[i] + C.gcur[i]; sgn -= gi * d[i]; gn += gi * gi; dn2 += d[i] * d[i]; }
            sgn = block_sum(C, sgn); gn = block_sum(C, gn); dn2 = block_sum(C, dn2);
            if (fabs(sgn) <= FLAT_TOL * sqrt(gn) * sqrt(dn2)) {
                // flat direction (gradient orthogonal to d within round-off): canonical sense, see qp_oracle.c
                get_row(C, c_out, C.t2);
                double an = 0;
                sgn = 0;
                K3R_LOOP for (int i = threadIdx.x; i < n; i += NT) { sgn += C.t2[i] * d[i]; an += C.t2[i] * C.t2[i]; }
                sgn = block_sum(C, sgn); an = block_sum(C, an);
                if (side_out > 0) sgn = -sgn;
                if (fabs(sgn) <= FLAT_TOL * sqrt(an) * sqrt(dn2)) {
                    double dmx = 0;
                    K3R_LOOP for (int i = threadIdx.x; i < n; i += NT) dmx = fmax(dmx, fabs(d[i]));
                    dmx = block_max(C, dmx);
                    int i0 = RT_NONE;
                    K3R_LOOP for (int i = threadIdx.x; i < n; i += NT) if (fabs(d[i]) > 0.25 * dmx) i0 = min(i0, i);
                    i0 = block_imin(C, i0);
                    sgn = (i0 == RT_NONE) ? 1.0 : d[i0];
                }
            }
        }
        if (sgn < 0) { __syncthreads(); for (int i = threadIdx.x; i < n; i += NT) d[i] = -d[i]; __syncthreads(); }
        Amul(C, d, C.Adx);
        double dn = 0, xmax = 0;
        K3R_LOOP for (int i = threadIdx.x; i < n; i += NT) { dn += d[i] * d[i]; xmax = fmax(xmax, fabs(C.x[i])); }
        dn = sqrt(block_sum(C, dn));
        xmax = block_max(C, xmax);
        const double pdel = C.tie * (1.0 + xmax);
        double tmax = INFINITY;
        K3R_LOOP for (int cc = threadIdx.x; cc < mt; cc += NT) {
            if (C.st[cc] != 0 || C.anorm[cc] == 0.0f) continue;
            int sd; double num, den;
            if (rt_exch(C, cc, dn, sd, num, den)) tmax = fmin(tmax, (num + pdel * C.anorm[cc]) * (1.0 / den));
        }
        tmax = block_min(C, tmax);
        if (!(tmax < INFINITY)) return EEPACC_QP_UNBOUNDED;
        int key = RT_NONE;
        K3R_LOOP for (int cc = threadIdx.x; cc < mt; cc += NT) {
            if (C.st[cc] != 0 || C.anorm[cc] == 0.0f) continue;
            int sd; double num, den;
            if (rt_exch(C, cc, dn, sd, num, den) && num * (1.0 / den) <= tmax) key = min(key, 2 * cc + sd);
        }
        key = block_imin(C, key);
        double tmin;
        {
            int sd; double num, den;
            rt_exch(C, key >> 1, dn, sd, num, den);
            tmin = (num <= pdel * C.anorm[key >> 1]) ? 0.0 : num / den;
        }
        const int j = key >> 1, jside = (key & 1) ? 1 : -1;
        K3R_LOOP for (int i = threadIdx.x; i < n; i += NT) C.x[i] += tmin * d[i];
        K3R_LOOP for (int cc = threadIdx.x; cc < mt; cc += NT) C.Ax[cc] += tmin * C.Adx[cc];
        __syncthreads();
        if (threadIdx.x == 0) { if (jside < 0) C.lbc[j] = C.Ax[j]; else C.ubc[j] = C.Ax[j]; }
        const double an2 = get_rowF(C, j, C.gF);
        givens_add(C, j, jside, C.gF, an2, kpd + 1, kpd, false);
        if (threadIdx.x == 0) { C.yW[kpd] = 0.0; if (zero_step) C.tabu[j] = 1; }
        __syncthreads();
        n_exch++;
        first = false; moved = true;
    }
    if (moved) reanchor(C, true);
    return 0;
}

// Warm start: rebuild the factorisation for the kept working set under the NEW data.
// Unit rows -> fixed variables (X list, stored order kept, a second unit row on the same variable
// is dropped); general rows -> Householder QR of A_gF' (nF x nG):  A_gF' = Qh [Rh; 0],
// Qf = Qh with reversed columns, T(r, c) = Rh(nF-1-c, nF-1-r).  Rows whose new bound is infinite
// or that became linearly dependent are dropped.  Returns 1 if nF exceeds the capacity.
template <class CT> K3R_FN __device__ int warm_setup(CT& C, int old_nA, const short* gidx, const signed char* gside, const double* gy)
{
    const int n = C.n, mt = C.mt, m = C.m;
    K3R_LOOP for (int c = threadIdx.x; c < mt; c += NT) C.st[c] = 0;
    short* cand = reinterpret_cast<short*>(C.t1);        // [n] constraint index of general candidates
    short* tmpc = reinterpret_cast<short*>(C.t3);        // classification scratch: [3n] shorts
    double* tmpv = C.t2;                                 // [n] unit coefficient
    double* tmpy = C.dx;                                 // [n] stored dual, by stored position
    double* ycand = C.t4;                                // [n] dual of general candidates
    double* v0a = C.csA;                                 // v0 of every reflector
    double* beta = C.snA;                                // 2 / v'v
    double* rdiag = C.sub;                               // Rh(k,k)
    double* cn2 = C.dbw;                                 // squared norm of the original row
    __syncthreads();
    // 1. classify the stored working set in parallel (position k was stored at index n-1-k)
    K3R_LOOP for (int k = threadIdx.x; k < old_nA; k += NT) {
        const int r = n - 1 - k;
        const int c = gidx[r], sd = gside[r];
        const double bnd = sd < 0 ? lb_new(C, c) : ub_new(C, c);
        int jv = -1; double cv = 0.0; const int ok = !isinf(bnd);
        if (ok) {
            const RowW rd = row_load(C, c);
            if (!s_unit(C, rd, jv, cv)) jv = -1;
        }
        tmpc[3 * k] = (short)(ok ? c : -1); tmpc[3 * k + 1] = (short)jv; tmpc[3 * k + 2] = (short)sd;
        tmpv[k] = cv;
        tmpy[k] = gy[r];                                  // coalesced: thread 0 below never touches global memory
    }
    K3R_LOOP for (int j = threadIdx.x; j < n; j += NT) C.fpos[j] = 0;
    __syncthreads();
#ifndef EEP_HOST_EMU
    // X list, general candidates and free list by ordered stream compaction in ONE WARP (ballots + popcounts; the
    // order of the stored working set is kept, exactly what the serial statement below does).  A second unit row
    // on a variable that is already pinned is dropped: the first occurrence is found with a shared atomicMin.
    int* first = reinterpret_cast<int*>(C.t1) + n;       // [n] (cand uses the first n shorts of t1)
    K3R_LOOP for (int j = threadIdx.x; j < n; j += NT) first[j] = 0x7fffffff;
    __syncthreads();
    K3R_LOOP for (int k = threadIdx.x; k < old_nA; k += NT)
        if (tmpc[3 * k] >= 0 && tmpc[3 * k + 1] >= 0) atomicMin(&first[tmpc[3 * k + 1]], k);
    __syncthreads();
    if (threadIdx.x < 32) {
        const unsigned lane = threadIdx.x, lt = (1u << lane) - 1u;
        int nx = 0, ng = 0;
        for (int k0 = 0; k0 < old_nA; k0 += 32) {
            const int k = k0 + (int)lane;
            int c = -1, jv = -1;
            if (k < old_nA) { c = tmpc[3 * k]; jv = tmpc[3 * k + 1]; }
            const bool isX = c >= 0 && jv >= 0 && first[jv] == k;
            const bool isG = c >= 0 && jv < 0;
            const unsigned mx = __ballot_sync(0xffffffffu, isX), mg = __ballot_sync(0xffffffffu, isG);
            if (isX) {
                const int q = nx + __popc(mx & lt);
                C.xvar[q] = (short)jv; C.xcon[q] = (short)c; C.xcv[q] = tmpv[k]; C.yX[q] = tmpy[k];
                C.fpos[jv] = (short)(-1 - q);
                C.st[c] = (signed char)tmpc[3 * k + 2];
            } else if (isG) {
                const int q = ng + __popc(mg & lt);
                cand[q] = (short)c; ycand[q] = tmpy[k]; C.st[c] = (signed char)tmpc[3 * k + 2];
            }
            nx += __popc(mx); ng += __popc(mg);
        }
        __syncwarp();
        int nf = 0;
        for (int j0 = 0; j0 < n; j0 += 32) {
            const int j = j0 + (int)lane;
            const bool fr = j < n && C.fpos[j] >= 0;
            const unsigned mf = __ballot_sync(0xffffffffu, fr);
            if (fr) {
                const int q = nf + __popc(mf & lt);
                if (q < C.fc) C.flist[q] = (short)j;
                C.fpos[j] = (short)q;
            }
            nf += __popc(mf);
        }
        if (lane == 0) { C.ired[4] = ng; C.ired[5] = nx; C.ired[6] = nf; }
    }
#else
    // thread 0: X list, general candidates, free list
    if (threadIdx.x == 0) {
        int nx = 0, ng = 0;
        for (int k = 0; k < old_nA; ++k) {
            const int c = tmpc[3 * k], jv = tmpc[3 * k + 1];
            if (c < 0) continue;
            if (jv >= 0) {
                if (C.fpos[jv] < 0) continue;             // coordinate already pinned: dependent, drop
                C.xvar[nx] = (short)jv; C.xcon[nx] = (short)c; C.xcv[nx] = tmpv[k]; C.yX[nx] = tmpy[k];
                C.fpos[jv] = (short)(-1 - nx);
                C.st[c] = (signed char)tmpc[3 * k + 2];
                ++nx;
            } else {
                cand[ng] = (short)c; ycand[ng] = tmpy[k]; C.st[c] = (signed char)tmpc[3 * k + 2];
                ++ng;
            }
        }
        int nf = 0;
        for (int j = 0; j < n; ++j)
            if (C.fpos[j] >= 0) { if (nf < C.fc) C.flist[nf] = (short)j; C.fpos[j] = (short)nf; ++nf; }
        C.ired[4] = ng; C.ired[5] = nx; C.ired[6] = nf;
    }
#endif
    __syncthreads();
    int cnt = C.ired[4];
    const int nX = C.ired[5], nF = C.ired[6];
    PH(1);
    C.nX = nX; C.nF = nF;
    if (nF > C.fc) return 1;
    if (cnt > nF) {                                       // more general rows than free variables: the tail is dependent
        __syncthreads();
        K3R_LOOP for (int k = nF + threadIdx.x; k < cnt; k += NT) C.st[cand[k]] = 0;
        cnt = nF;
        __syncthreads();
    }
    // 2. stage the general rows (free part) as columns of W (= Qf storage): LPC lanes per row
    for (int k0 = 0; k0 < cnt; k0 += NT / LPC) {          // warp-uniform trip count (full-mask shuffles inside)
        const int k = k0 + threadIdx.x / LPC, lane = threadIdx.x % LPC;
        double s2 = 0;
        if (k < cnt) {
            const RowW rd = row_load(C, cand[k]);
            for (int j = lane; j < n; j += LPC) {
                const double v = s_entry(C, rd, j);
                s2 += v * v;
                const int f = C.fpos[j];
                if (f >= 0) QF(f, k) = v;
            }
        }
        s2 = group_sum(s2);
        if (k < cnt && lane == 0) cn2[k] = s2;
    }
    __syncthreads();
    PH(2);
    // 3. Householder QR; a dependent column is swapped with the last and dropped.  LPC lanes share a
    //    column; the lane group of column k+1 also hands the next pivot norm over (C.red[6 + slot]), so a
    //    regular column costs ONE barrier.
    int k = 0, slot = 0;
    bool have_s2 = false;
    double s2 = 0;
    while (k < cnt) {
        if (!have_s2) {
            double a2 = 0;
            K3R_LOOP for (int i = k + threadIdx.x; i < nF; i += NT) a2 += QF(i, k) * QF(i, k);
            s2 = block_sum(C, a2);
        }
        have_s2 = false;
        const double ref = cn2[k] > 1.0 ? cn2[k] : 1.0;
        if (!(s2 > EPS_LI * EPS_LI * ref)) {
            const int last = cnt - 1;
            __syncthreads();
            if (threadIdx.x == 0) C.st[cand[k]] = 0;
            __syncthreads();
            if (k != last) {
                K3R_LOOP for (int i = threadIdx.x; i < nF; i += NT) QF(i, k) = QF(i, last);
                if (threadIdx.x == 0) { cand[k] = cand[last]; ycand[k] = ycand[last]; cn2[k] = cn2[last]; }
            }
            cnt = last;
            __syncthreads();
            continue;
        }
        const double sigma = sqrt(s2), x0 = QF(k, k);
        const double alpha = (x0 >= 0.0) ? -sigma : sigma;
        const double v0 = x0 - alpha;
        const double bt = 1.0 / (s2 - x0 * alpha);
#ifndef EEP_HOST_EMU
        // lanes per column: as many as the remaining columns leave room for (4, 8 or 16): the per-lane chain of the
        // dot product and of the update is what a column step costs
        const int ncol = cnt - k - 1;
        const int lpc = (ncol * 16 <= NT) ? 16 : (ncol * 8 <= NT) ? 8 : 4;
        for (int j0 = k + 1; j0 < cnt; j0 += NT / lpc) {
            const int j = j0 + (int)threadIdx.x / lpc, lane = threadIdx.x % lpc;
            const bool act = j < cnt;
            double d = 0, d1 = 0;
            if (act) {
                int i = k + 1 + lane;
                for (; i + lpc < nF; i += 2 * lpc) { d += QF(i, k) * QF(i, j); d1 += QF(i + lpc, k) * QF(i + lpc, j); }
                if (i < nF) d += QF(i, k) * QF(i, j);
            }
            if (act && lane == 0) d += v0 * QF(k, j);     // consumed before the shuffle: lane 0 rewrites it below
            d += d1;
            d += __shfl_xor_sync(0xffffffffu, d, 1); d += __shfl_xor_sync(0xffffffffu, d, 2);
            if (lpc >= 8) d += __shfl_xor_sync(0xffffffffu, d, 4);
            if (lpc >= 16) d += __shfl_xor_sync(0xffffffffu, d, 8);
            double nn = 0;
            if (act) {
                d *= bt;
                for (int i = k + 1 + lane; i < nF; i += lpc) {
                    const double v = QF(i, j) - d * QF(i, k);
                    QF(i, j) = v;
                    nn += v * v;
                }
            }
            if (j0 == k + 1) {                            // warp-uniform: first pass only
                nn += __shfl_xor_sync(0xffffffffu, nn, 1); nn += __shfl_xor_sync(0xffffffffu, nn, 2);
                if (lpc >= 8) nn += __shfl_xor_sync(0xffffffffu, nn, 4);
                if (lpc >= 16) nn += __shfl_xor_sync(0xffffffffu, nn, 8);
                if (act && j == k + 1 && lane == 0) C.red[6 + slot] = nn;
            }
            if (act && lane == 0) QF(k, j) -= d * v0;
        }
#else
        for (int j0 = k + 1; j0 < cnt; j0 += NT / LPC) {
            const int j = j0 + threadIdx.x / LPC, lane = threadIdx.x % LPC;
            const bool act = j < cnt;
            double d = 0, d1 = 0;
            if (act) {
                int i = k + 1 + lane;
                for (; i + LPC < nF; i += 2 * LPC) { d += QF(i, k) * QF(i, j); d1 += QF(i + LPC, k) * QF(i + LPC, j); }
                if (i < nF) d += QF(i, k) * QF(i, j);
            }
            if (act && lane == 0) d += v0 * QF(k, j);     // consumed before the shuffle: lane 0 rewrites it below
            d = group_sum(d + d1);
            double nn = 0;
            if (act) {
                d *= bt;
                for (int i = k + 1 + lane; i < nF; i += LPC) {
                    const double v = QF(i, j) - d * QF(i, k);
                    QF(i, j) = v;
                    nn += v * v;
                }
            }
            if (j0 == k + 1) {                            // warp-uniform: first pass only
                nn = group_sum(nn);
                if (act && j == k + 1 && lane == 0) C.red[6 + slot] = nn;
            }
            if (act && lane == 0) QF(k, j) -= d * v0;
        }
#endif
        if (threadIdx.x == 0) { v0a[k] = v0; beta[k] = bt; rdiag[k] = alpha; }
        __syncthreads();
        ++k;
        if (k < cnt) { s2 = C.red[6 + slot]; have_s2 = true; slot ^= 1; }
    }
    PH(3);
    const int nAC = cnt, nZ = nF - nAC;
    // 4. T block of M from Rh:  M(r, c) = Rh(nF-1-c, nF-1-r),  c >= r >= nZ
    K3R_LOOP for (int r = nZ + threadIdx.x; r < nF; r += NT) {
        const int b = nF - 1 - r;
        for (int c = r; c < nF; ++c) {
            const int a = nF - 1 - c;
            MM(r, c) = (a == b) ? rdiag[a] : QF(a, b);
        }
    }
    __syncthreads();
    // 5. form Qh = H_0 H_1 ... H_(nAC-1).
#ifndef EEP_HOST_EMU
    // Column j of Qh is H_0 .. H_min(j, nAC-1) e_j: the columns are INDEPENDENT once the reflectors are read-only.  They
    // are copied (with their leading entry made explicit) into the scratch that is dead during the set-up -- the
    // ratio-test slots and, at the first-level capacity, the homotopy state behind them -- and every column is then
    // built by a pair of lanes with no block barrier at all (the in-place backward accumulation below pays one barrier
    // and a full pass over the trailing block per reflector: 59-70 k cycles of a ~450-700 k cycle solve).  The column
    // lands straight in its final, reversed position (step 6 disappears).
    {
        const int navail = (2 * ((mt + NT - 1) / NT) + (n + NT - 1) / NT) * NT + (C.hs_sm ? 4 * mt : 0);
        if (nAC > 0 && nF * nAC <= navail) {
            double* Vs = C.tt;                                   // [nAC][nF]: reflector k = Vs[k * nF + (k .. nF-1)]
            K3R_LOOP for (int e = threadIdx.x; e < nAC * nF; e += NT) {
                const int k = e / nF, i = e - k * nF;
                Vs[e] = (i > k) ? QF(i, k) : (i == k ? v0a[k] : 0.0);
            }
            __syncthreads();
            for (int j0 = 0; j0 < nF; j0 += NT / 2) {            // warp-uniform trip count (full-mask shuffles inside)
                const int j = j0 + (int)(threadIdx.x >> 1), lane = threadIdx.x & 1;
                const bool act = j < nF;
                const int dc = act ? nF - 1 - j : 0;             // Qf = Qh with reversed columns
                if (act) for (int i = lane; i < nF; i += 2) QF(i, dc) = (i == j) ? 1.0 : 0.0;
                __syncwarp();
                for (int kk = nAC - 1; kk >= 0; --kk) {
                    const bool on = act && kk <= j;              // H_kk e_j = e_j for kk > j
                    const double* v = Vs + kk * nF;
                    const int i0 = kk + ((lane - kk) & 1);       // first row >= kk of this lane's parity
                    double d = 0;
                    if (on) for (int i = i0; i < nF; i += 2) d += v[i] * QF(i, dc);
                    d += __shfl_xor_sync(0xffffffffu, d, 1);
                    d *= beta[kk];
                    if (on) for (int i = i0; i < nF; i += 2) QF(i, dc) -= d * v[i];
                    __syncwarp();
                }
            }
            K3R_LOOP for (int kq = threadIdx.x; kq < nAC; kq += NT) { C.yW[nF - 1 - kq] = ycand[kq]; C.widx[nF - 1 - kq] = cand[kq]; }
            C.nZ = nZ;
            __syncthreads();
            PH(4);
            return 0;
        }
    }
#endif
    K3R_LOOP for (int j = nAC + threadIdx.x; j < nF; j += NT)
        for (int i = 0; i < nF; ++i) QF(i, j) = (i == j) ? 1.0 : 0.0;
    __syncthreads();
    // Reflector kk is applied to the columns right of it by LPC lanes per column.  Column kk+1 still holds
    // the raw reflector kk+1 at that point: its lane group turns it into H_(kk+1) e_(kk+1) on the fly
    // (rows above the diagonal become zero), so one barrier per reflector is enough.
    for (int kk = nAC - 1; kk >= 0; --kk) {
        const double v0 = v0a[kk], bt = beta[kk];
        for (int j0 = kk + 1; j0 < nF; j0 += NT / LPC) {
            const int j = j0 + threadIdx.x / LPC, lane = threadIdx.x % LPC;
            const bool act = j < nF;
            const bool conv = act && j == kk + 1 && j < nAC;
            const double cj = conv ? -beta[j] * v0a[j] : 1.0;      // scale of the raw entries below the diagonal
            const double dj = conv ? 1.0 - beta[j] * v0a[j] * v0a[j] : 0.0;
            double d = 0;
            if (act) {
                for (int i = kk + 1 + lane; i < nF; i += LPC) {
                    const double c = (conv && i == j) ? dj : cj * QF(i, j);
                    d += QF(i, kk) * c;
                }
                if (lane == 0 && !conv) d += v0 * QF(kk, j);
            }
            d = group_sum(d) * bt;
            if (act) {
                for (int i = kk + 1 + lane; i < nF; i += LPC) {
                    const double c = (conv && i == j) ? dj : cj * QF(i, j);
                    QF(i, j) = c - d * QF(i, kk);
                }
                if (lane == 0) QF(kk, j) = (conv ? 0.0 : QF(kk, j)) - d * v0;
                if (conv) for (int i = lane; i < kk; i += LPC) QF(i, j) = 0.0;
            }
        }
        __syncthreads();
    }
    if (nAC > 0) {
        const double v0 = v0a[0], bt = beta[0];
        K3R_LOOP for (int i = threadIdx.x; i < nF; i += NT) {
            const double vi = QF(i, 0);
            QF(i, 0) = (i == 0) ? 1.0 - bt * v0 * v0 : -bt * v0 * vi;
        }
        __syncthreads();
    }
    // 6. reverse the columns, 7. working-set bookkeeping by Qf column
    K3R_LOOP for (int e = threadIdx.x; e < nF * (nF / 2); e += NT) {
        const int c = e / nF, i = e - c * nF;
        const double a = QF(i, c), b2 = QF(i, nF - 1 - c);
        QF(i, c) = b2; QF(i, nF - 1 - c) = a;
    }
    K3R_LOOP for (int kq = threadIdx.x; kq < nAC; kq += NT) { C.yW[nF - 1 - kq] = ycand[kq]; C.widx[nF - 1 - kq] = cand[kq]; }
    C.nZ = nZ;
    __syncthreads();
    PH(4);
    return 0;
}

// all variables fixed at their (relaxed) lower far bound; position k <-> variable n-1-k (k3_qp.cu order)
template <class CT> K3R_FN __device__ void cold_ws(CT& C)
{
    const int n = C.n, m = C.m, mt = C.mt;
    __syncthreads();
    K3R_LOOP for (int c = threadIdx.x; c < mt; c += NT) C.st[c] = 0;
    __syncthreads();
    K3R_LOOP for (int k = threadIdx.x; k < n; k += NT) {
        const int v = n - 1 - k;
        C.xvar[k] = (short)v; C.xcon[k] = (short)(m + v); C.xcv[k] = 1.0;
        C.fpos[v] = (short)(-1 - k);
        C.st[m + v] = -1;
        C.yX[k] = 0.5 + 0.5 * (double)(m + v) / (double)mt;   // distinct initial duals ("ramping")
    }
    C.nX = n; C.nF = 0; C.nZ = 0;
    __syncthreads();
}

// auxiliary QP for which the current (x, y, working set) is optimal: relaxed bounds + anchored gradient
template <class CT> K3R_FN __device__ void setup_aux(CT& C)
{
    const int mt = C.mt;
    Amul(C, C.x, C.Ax);
    K3R_LOOP for (int c = threadIdx.x; c < mt; c += NT) {
        int s = C.st[c];
        double ax = C.Ax[c], ln = lb_new(C, c), un = ub_new(C, c);
        bool eq = ln == un;
        if (s < 0) { C.lbc[c] = ax; C.ubc[c] = eq ? ax : (isinf(un) ? INFINITY : ax + RELAX); }
        else if (s > 0) { C.ubc[c] = ax; C.lbc[c] = eq ? ax : (isinf(ln) ? -INFINITY : ax - RELAX); }
        else { C.lbc[c] = isinf(ln) ? -INFINITY : ax - RELAX; C.ubc[c] = isinf(un) ? INFINITY : ax + RELAX; }
    }
    __syncthreads();
    reanchor(C, false);
}

}  // namespace k3r

extern "C" __host__ __device__ size_t eep_qpr_smem_bytes(int n, int m, int fc);
__host__ __device__ size_t eep_qpr_smem_base(int n, int m, int fc);
// the homotopy state (4 (m + n) doubles) lives in shared memory whenever two CTAs per SM still fit with it
__host__ __device__ inline bool eep_qpr_hs_in_smem(int n, int m, int fc)
{
    return eep_qpr_smem_base(n, m, fc) + 32 * (size_t)(n + m) <= (size_t)113 * 1024;
}
#ifndef EEP_HOST_EMU
extern __shared__ __align__(16) unsigned char k3r_smem[];
#endif

// ---- row table builder (once per set-up; also runs on the host in the emulation build) ----
// consts[96] = cvals[32] (control coefficients, [0] = 1), cva[32], cvb[32] (coefficient on v_k = a + b v_est(k)).
// Returns 0, or a positive code if a row does not fit the encoding (set-up then fails loudly).
__device__ int k3r_build_rowtab(int ctrl, const eepacc_params& p, int m, unsigned long long* tab, double* consts)
{
    const int NH = EEP_N, nu = (ctrl == EEPACC_FB) ? 6 : (ctrl == EEPACC_AB) ? 5 : 2;
    double* cvals = consts; double* cva = consts + 32; double* cvb = consts + 64;
    for (int i = 0; i < 96; ++i) consts[i] = 0.0;
    int ncv = 1, npair = 1;
    cvals[0] = 1.0;
    double ve0[NH + 1], ve1[NH + 1], ve2[NH + 1];
    const double big = 1048576.0;
    for (int i = 0; i <= NH; ++i) { ve0[i] = 0.0; ve1[i] = big; ve2[i] = 3.0; }
    k3r::VestSrc s0{ ve0 }, s1{ ve1 }, s2{ ve2 };
    for (int c = 0; c < m; ++c) {
        RowDesc r0, r1, r2;
        eep_row<false>(ctrl, p, 1.0, s0, c, r0);
        eep_row<false>(ctrl, p, 1.0, s1, c, r1);
        eep_row<false>(ctrl, p, 1.0, s2, c, r2);
        if (r0.kst != r1.kst || r0.nctl != r1.nctl || r0.cS != r1.cS) return 1;
        if (r0.cv0 != r1.cv0 || r0.cv1 != r1.cv1 || r0.cv2 != r1.cv2) return 2;          // control coefficients must not depend on v_est
        if (!(r0.cS == 0.0 || r0.cS == 1.0)) return 3;
        if (r0.kst < -2 || r0.kst > NH) return 4;
        int aux = 0;
        if (r0.kst >= 0) {
            const double a = r0.cV, b = (r1.cV - r0.cV) / big;
            if (fabs(r2.cV - (a + b * 3.0)) > 1e-12 * (1.0 + fabs(r2.cV))) return 5;      // must be affine in v_est(k)
            int id = -1;
            for (int q = 0; q < npair; ++q) if (cva[q] == a && cvb[q] == b) { id = q; break; }
            if (id < 0) { if (npair >= 32) return 6; id = npair++; cva[id] = a; cvb[id] = b; }
            aux = id | (r0.cS != 0.0 ? 128 : 0);
        } else if (r0.kst == -2) {
            aux = r0.cj0 / nu - 1;
            if (aux < 0 || aux >= NH || r0.cV != 0.0 || r0.cS != 0.0) return 7;
        } else if (r0.cV != 0.0 || r0.cS != 0.0) return 8;
        const int cj[3] = { r0.cj0, r0.cj1, r0.cj2 };
        const double cv[3] = { r0.cv0, r0.cv1, r0.cv2 };
        unsigned long long w = (unsigned long long)(unsigned char)(signed char)r0.kst | ((unsigned long long)aux << 8);
        for (int e = 0; e < 3; ++e) {
            int jj = 255, id = 0;
            if (e < r0.nctl) {
                jj = cj[e];
                if (jj < 0 || jj >= 255) return 9;
                id = -1;
                for (int q = 0; q < ncv; ++q) if (cvals[q] == cv[e]) { id = q; break; }
                if (id < 0) { if (ncv >= 32) return 10; id = ncv++; cvals[id] = cv[e]; }
            }
            w |= (unsigned long long)jj << (16 + 8 * e);
            w |= (unsigned long long)id << (40 + 8 * e);
        }
        tab[c] = w;
    }
    return 0;
}

// Lay the solver's work arrays out over `base` (shared memory, or a global scratch block for a hand-over) for
// capacity fc.  Writes nothing.
template <int NU, int NB>
__device__ __forceinline__ void k3r_setup(k3r::CtxT<NU, NB>& C, const eep_qp_io& io, unsigned char* base, const int fc)
{
    const int n = io.n, m = io.m, mt = n + m;
    C.n = n; C.m = m; C.mt = mt; C.fc = fc; C.ldq = (fc + 1) | 1;
    double* gS;
    {
        double* p = reinterpret_cast<double*>(base);
        C.Q = p; p += (size_t)(fc + 1) * C.ldq;
        C.M = p; p += (size_t)fc * (fc + 1) / 2 + fc + 2;
        C.sub = p; p += fc + 2;
        C.x = p; p += n; C.yX = p; p += n; C.xcv = p; p += n; gS = p; p += n;
        // the next nine vectors are contiguous: refactor_R borrows them as scratch (nothing in them is live then)
        C.gcur = p; p += n; C.dx = p; p += n; C.dg = p; p += n;
        C.t1 = p; p += n; C.t2 = p; p += n; C.t3 = p; p += n; C.t4 = p; p += n;
        C.dyX = p; p += n; C.dbX = p; p += n;
        C.yW = p; p += fc + 1; C.dyW = p; p += fc + 1; C.dbw = p; p += fc + 1; C.wv = p; p += fc + 1;
        C.csA = p; p += fc + 1; C.snA = p; p += fc + 1; C.pF = p; p += fc + 1; C.uF = p; p += fc + 1;
        C.tF = p; p += fc + 1; C.gF = p; p += fc + 1;
        C.anorm = reinterpret_cast<float*>(p); p += (mt + 1) / 2;
        const int NH = EEP_N;
        C.PS = p; p += (NH + 1) * NH; C.PV = p; p += (NH + 1) * NH; C.vest = p; p += NH + 1;
        C.Dc = p; p += NH * 9; C.Cc = p; p += NH * 9; C.Dg = p; p += NH * 4;
        C.sv = p; p += NH + 1; C.vv = p; p += NH + 1; C.drv = p; p += NH; C.yc = p; p += NH * 3;
        C.aS = p; p += NH + 1; C.aV = p; p += NH + 1; C.aD = p; p += NH;
        C.rS = p; p += fc + 1; C.rV = p; p += fc + 1; C.rv0 = p; p += fc + 1; C.rv1 = p; p += fc + 1; C.rv2 = p; p += fc + 1;
        C.red = p; p += 16;
        C.ired = reinterpret_cast<int*>(p); p += 4;      // 8 ints
        C.tt = p; p += (size_t)(2 * ((mt + NT - 1) / NT) + (n + NT - 1) / NT) * NT;
        C.hs_sm = nullptr;
        if (eep_qpr_hs_in_smem(n, m, fc)) { C.hs_sm = p; p += 4 * (size_t)mt; }
        double* cst = p; p += 96;
        C.g = gS; C.gw = gS; C.rtab = io.rowtab; C.cvals = cst; C.cva = cst + 32; C.cvb = cst + 64;
        short* sp = reinterpret_cast<short*>(p);
        C.rk = sp; sp += fc + 1; C.rj0 = sp; sp += fc + 1; C.rj1 = sp; sp += fc + 1; C.rj2 = sp; sp += fc + 1; C.rj3 = sp; sp += fc + 1;
        C.flist = sp; sp += fc + 1; C.widx = sp; sp += fc + 1;
        C.fpos = sp; sp += n; C.xvar = sp; sp += n; C.xcon = sp; sp += n;
        C.st = reinterpret_cast<signed char*>(sp);
        C.tabu = C.st + mt;
#ifdef EEP_K3_TIMERS
        C.ph = reinterpret_cast<long long*>(base + ((eep_qpr_smem_bytes(n, m, fc) - 34 * 8) & ~(size_t)7));
#endif
    }
}
// per-controller coefficient tables: the same for every vehicle, but they live in the work area (re-load them
// whenever something else used it)
template <class CT> __device__ __forceinline__ void k3r_tables(CT& C, const eep_qp_io& io)
{
    double* cst = const_cast<double*>(C.cvals);
    K3R_LOOP for (int e = threadIdx.x; e < 96; e += NT) cst[e] = io.rowc[e];
    if (threadIdx.x == 0) C.vest[EEP_N] = 0.0;
}
// One vehicle's QP.  Returns the status written to io.status, or K3R_DEFER (nothing written, warm state
// untouched) when the free variables exceed the capacity.
template <class CT>
K3R_FN __device__ int k3r_vehicle(CT& C, const eep_qp_io& io, const long long veh)
{
    using namespace k3r;
    const int n = C.n, m = C.m, mt = C.mt;
    __syncthreads();
    int status = EEPACC_QP_OK, iter = 0;
    const int max_iter = 5 * (n + m);
    const int restart_after = (4 * n > 400 ? 4 * n : 400);
    unsigned long long n_exch = 0, n_refac = 0;
#ifdef EEP_K3_TIMERS
    if (threadIdx.x == 0) { for (int q = 0; q < 32; ++q) C.ph[q] = 0; C.ph[32] = clock64(); }
#endif

    // gradient, bounds, structured block, Hessian scale, LP regularisation, row norms
    {
        const double* gg = io.g + veh * (size_t)n;
        const double* la = io.lba + veh * (size_t)m;
        const double* ua = io.uba + veh * (size_t)m;
        K3R_LOOP for (int i = threadIdx.x; i < n; i += NT) C.gw[i] = gg[i];
        C.lba = la; C.uba = ua;
        {
            // homotopy state over all m + n rows: in shared memory at the first-level capacity (registers, not shared
            // memory, hold the kernel to two CTAs per SM, so the 20 KB are free; ncu had 23 % of these accesses miss
            // L1), in the per-vehicle global scratch for the large-capacity hand-over launches
            double* hs = C.hs_sm ? C.hs_sm : io.hscr + veh * (size_t)(4 * mt);
            C.Ax = hs; C.lbc = hs + mt; C.ubc = hs + 2 * mt; C.Adx = hs + 3 * mt;
        }
        const double* sb = io.sblk + veh * (size_t)EEP_SBLK;
        K3R_LOOP for (int e = threadIdx.x; e < (EEP_N + 1) * EEP_N; e += NT) { C.PS[e] = sb[EEP_SB_PS + e]; C.PV[e] = sb[EEP_SB_PV + e]; }
        K3R_LOOP for (int e = threadIdx.x; e < EEP_N * 9; e += NT) { C.Dc[e] = sb[EEP_SB_DC + e]; C.Cc[e] = sb[EEP_SB_CC + e]; }
        K3R_LOOP for (int e = threadIdx.x; e < EEP_N * 4; e += NT) C.Dg[e] = sb[EEP_SB_DG + e];
        K3R_LOOP for (int e = threadIdx.x; e < EEP_N; e += NT) C.vest[e] = sb[EEP_SB_VEST + e];
        C.reg = (sb[EEP_SB_HMAX] == 0.0) ? LP_REG : 0.0;
        C.hscale = fmax(sb[EEP_SB_HSCALE], LP_REG);
        C.tie = (C.reg != 0.0) ? 0.0 : TIE_TOL;
        __syncthreads();
        K3R_LOOP for (int r = threadIdx.x; r < m; r += NT) {
            const RowW rd = row_load(C, r);
            double sa = 0;
            if (rd.kst >= 0) {
                const double cV = row_cV(C, rd), cS = (rd.aux & 128) ? 1.0 : 0.0;
                for (int i = 0; i < rd.kst && i < EEP_N; ++i)
                    sa += CT::nb * fabs(cS * C.PS[rd.kst * EEP_N + i] + cV * C.PV[rd.kst * EEP_N + i]);
            }
            if (rd.j0 != 255) sa += fabs(C.cvals[rd.i0]);
            if (rd.j1 != 255) sa += fabs(C.cvals[rd.i1]);
            if (rd.j2 != 255) sa += fabs(C.cvals[rd.i2]);
            if (rd.kst == -2) sa += CT::nb;
            C.anorm[r] = (float)sa;
        }
    }
    K3R_LOOP for (int j = threadIdx.x; j < n; j += NT) C.anorm[m + j] = 1.0f;
    K3R_LOOP for (int c = threadIdx.x; c < mt; c += NT) C.tabu[c] = 0;
    __syncthreads();
    PH(0);

    // ---------------- set-up ----------------
    const int was_warm = io.warm[veh];
    bool need_cold = !was_warm;
    K3R_LOOP for (int i = threadIdx.x; i < n; i += NT) C.x[i] = was_warm ? io.wx[veh * (size_t)n + i] : 0.0;
    if (!need_cold) {
        if (warm_setup(C, io.wnA[veh], io.widx + veh * (size_t)n, io.wside + veh * (size_t)n, io.wy + veh * (size_t)n))
            status = K3R_DEFER;
        else {
            n_refac++;
            if (refactor_R(C)) need_cold = true;
            PH(5);
        }
    }
    if (status == EEPACC_QP_OK) {
        if (need_cold) cold_ws(C);
        setup_aux(C);
        double ym = 1.0;
        K3R_LOOP for (int r = C.nZ + threadIdx.x; r < C.nF; r += NT) ym = fmax(ym, fabs(C.yW[r]));
        K3R_LOOP for (int k = threadIdx.x; k < C.nX; k += NT) ym = fmax(ym, fabs(C.yX[k]));
        C.ynoise = block_max(C, ym);
    }
    PH(6);

    // ---------------- homotopy ----------------
    while (status == EEPACC_QP_OK) {
        const int nZ = C.nZ, nF = C.nF, nX = C.nX;
        K3R_LOOP for (int i = threadIdx.x; i < n; i += NT) C.dg[i] = C.g[i] - C.gcur[i];
        K3R_LOOP for (int r = nZ + threadIdx.x; r < nF; r += NT) {
            int c = C.widx[r];
            double nb = C.st[c] < 0 ? lb_new(C, c) : ub_new(C, c);
            double cb = C.st[c] < 0 ? C.lbc[c] : C.ubc[c];
            C.dbw[r] = isinf(nb) ? 0.0 : nb - cb;
        }
        K3R_LOOP for (int k = threadIdx.x; k < nX; k += NT) {
            int c = C.xcon[k];
            double nb = C.st[c] < 0 ? lb_new(C, c) : ub_new(C, c);
            double cb = C.st[c] < 0 ? C.lbc[c] : C.ubc[c];
            C.dbX[k] = isinf(nb) ? 0.0 : nb - cb;
        }
        __syncthreads();
        PH(14);
        kkt_step(C, C.dg, C.dbw, C.dbX, C.dx, C.dyW, C.dyX);
        Amul(C, C.dx, C.Adx);
        PH(13);

        // ---- ratio tests, two-pass (qp_rules.cuh; oracle/qp_oracle.c:rt_select).  Pass 1 leaves every candidate's
        // ratio in the shared scratch C.tt (slot-major, one column per thread: no bank conflicts), pass 2 only
        // compares: no second evaluation, one division per candidate. ----
        double dxmax = 0, xmax = 0, dys = 0;
        K3R_LOOP for (int i = threadIdx.x; i < n; i += NT) { dxmax = fmax(dxmax, fabs(C.dx[i])); xmax = fmax(xmax, fabs(C.x[i])); }
        K3R_LOOP for (int r = nZ + threadIdx.x; r < nF; r += NT) dys = fmax(dys, fabs(C.dyW[r]));
        K3R_LOOP for (int k = threadIdx.x; k < nX; k += NT) dys = fmax(dys, fabs(C.dyX[k]));
        block_max3(C, dxmax, xmax, dys);
        const double thr = EPS_RATE * (1.0 + dys);
        const double pdel = C.tie * (1.0 + xmax), ddel = C.tie * C.ynoise;
        double tmax = INFINITY;
        double* tt = C.tt + threadIdx.x;
        int slot = 0;
        K3R_LOOP for (int c = threadIdx.x; c < mt; c += NT, slot += 2) {
            double t0 = INFINITY, t1 = INFINITY;
            if (C.st[c] == 0 && C.anorm[c] != 0.0f && C.tabu[c] != 2) {   // all-zero rows can never be added; tabu 2: left during this run of zero steps
                double num, den;
                if (rt_primal(C, c, 0, dxmax, num, den)) { const double inv = 1.0 / den; t0 = num * inv; tmax = fmin(tmax, (num + pdel * C.anorm[c]) * inv); }
                if (rt_primal(C, c, 1, dxmax, num, den)) { const double inv = 1.0 / den; t1 = num * inv; tmax = fmin(tmax, (num + pdel * C.anorm[c]) * inv); }
            }
            tt[slot * NT] = t0; tt[(slot + 1) * NT] = t1;
        }
        // dual candidates: X entries are positions 0..nX-1, the general row of Qf column r is position nX + nF-1-r
        const int dslot0 = 2 * ((mt + NT - 1) / NT);
        slot = dslot0;
        K3R_LOOP for (int q = threadIdx.x; q < nX + (nF - nZ); q += NT, ++slot) {
            int c; double yc, dy;
            if (q < nX) { c = C.xcon[q]; yc = C.yX[q]; dy = C.dyX[q]; }
            else { const int r = nF - 1 - (q - nX); c = C.widx[r]; yc = C.yW[r]; dy = C.dyW[r]; }
            double t0 = INFINITY, num, den;
            // equalities never leave; tabu 1: entered during this run of zero steps, stays
            if (lb_new(C, c) != ub_new(C, c) && C.tabu[c] != 1 && rt_dual(C.st[c], yc, dy, thr, ddel, num, den)) {
                const double inv = 1.0 / den; t0 = num * inv; tmax = fmin(tmax, (num + ddel) * inv);
            }
            tt[slot * NT] = t0;
        }
        tmax = block_min(C, tmax);
        // pass 2: lowest key among the candidates within tmax; packed = key * 128 + dual position
        int pk = RT_NONE;
        if (tmax < INFINITY) {
            slot = 0;
            K3R_LOOP for (int c = threadIdx.x; c < mt; c += NT, slot += 2) {
                if (tt[slot * NT] <= tmax) pk = min(pk, (2 * c) * 128);
                if (tt[(slot + 1) * NT] <= tmax) pk = min(pk, (2 * c + 1) * 128);
            }
            slot = dslot0;
            K3R_LOOP for (int q = threadIdx.x; q < nX + (nF - nZ); q += NT, ++slot) {
                if (tt[slot * NT] <= tmax) {
                    const int c = (q < nX) ? C.xcon[q] : C.widx[nF - 1 - (q - nX)];
                    pk = min(pk, (2 * mt + c) * 128 + q);
                }
            }
            pk = block_imin(C, pk);
        }
        const bool full = (pk == RT_NONE);
        double tau = 1.0;
        const int key = full ? RT_NONE : pk >> 7, qsel = pk & 127;
        if (!full) {
            double num = 0, den = 1, del;
            if (key < 2 * mt) { rt_primal(C, key >> 1, key & 1, dxmax, num, den); del = pdel * C.anorm[key >> 1]; }
            else {
                int c; double yc, dy;
                if (qsel < nX) { c = C.xcon[qsel]; yc = C.yX[qsel]; dy = C.dyX[qsel]; }
                else { const int r = nF - 1 - (qsel - nX); c = C.widx[r]; yc = C.yW[r]; dy = C.dyW[r]; }
                rt_dual(C.st[c], yc, dy, thr, ddel, num, den); del = ddel;
            }
            tau = (num <= del) ? 0.0 : num / den;       // within round-off of zero: a zero step
        }
        C.ynoise += tau * dys;
        // anti-cycling (see oracle/qp_oracle.c): constraints that enter while the homotopy makes no
        // progress may not leave through the dual test until it does
        const bool zero_step = !(tau > 1e-14);
        if (!zero_step) for (int c = threadIdx.x; c < mt; c += NT) C.tabu[c] = 0;

        // take the step
        K3R_LOOP for (int i = threadIdx.x; i < n; i += NT) {
            C.x[i] += tau * C.dx[i];
            C.gcur[i] = full ? C.g[i] : C.gcur[i] + tau * C.dg[i];
        }
        K3R_LOOP for (int r = nZ + threadIdx.x; r < nF; r += NT) C.yW[r] += tau * C.dyW[r];
        K3R_LOOP for (int k = threadIdx.x; k < nX; k += NT) C.yX[k] += tau * C.dyX[k];
        K3R_LOOP for (int c = threadIdx.x; c < mt; c += NT) {
            C.Ax[c] += tau * C.Adx[c];
            double ln = lb_new(C, c), un = ub_new(C, c);
            if (!isinf(ln)) C.lbc[c] = full ? ln : C.lbc[c] + tau * (ln - C.lbc[c]);
            if (!isinf(un)) C.ubc[c] = full ? un : C.ubc[c] + tau * (un - C.ubc[c]);
        }
        __syncthreads();
        PH(14);
        if (full) break;
        if (++iter > max_iter) { status = EEPACC_QP_MAXITER; break; }
        if (iter == restart_after) {
            // the warm-started path wanders (degenerate / non-convex stretches): start over from an
            // all-fixed working set at the current point -- the cold-start path
            cold_ws(C);
            setup_aux(C);
            K3R_LOOP for (int c = threadIdx.x; c < mt; c += NT) C.tabu[c] = 0;
            __syncthreads();
            continue;
        }
        if (io.dbg && veh == 0 && threadIdx.x == 0 && iter <= io.dbg_cap) {
            double* o = io.dbg + (size_t)(iter - 1) * 4;
            o[0] = tau; o[1] = (double)key; o[2] = (double)(C.nX + C.nF - C.nZ); o[3] = (double)n_exch;
        }

        if (key < 2 * mt) {
            // ---------- primal blocking: add constraint ----------
            const int blk = key >> 1, side = (key & 1) ? 1 : -1;
            if (threadIdx.x == 0) {
                if (side < 0) C.lbc[blk] = C.Ax[blk]; else C.ubc[blk] = C.Ax[blk];
                if (zero_step) C.tabu[blk] = 1;
            }
            get_row(C, blk, C.t4);                      // full row (the exchange needs its fixed part)
            double an2 = 0;
            K3R_LOOP for (int j = threadIdx.x; j < n; j += NT) an2 += C.t4[j] * C.t4[j];
            an2 = block_sum(C, an2);
            gatherF(C, C.t4);
            const int nZc = C.nZ;
            if (givens_add(C, blk, side, C.gF, an2, nZc, nZc, true) == 0) {
                if (threadIdx.x == 0) C.yW[nZc - 1] = 0.0;
                __syncthreads();
            } else {
                // linearly dependent: a = A_W' lam (general rows and fixed variables), exchange by a dual ratio test
                const int nFc = C.nF, nXc = C.nX;
                double* lamG = C.dyW;                   // by Qf column (dyW / dyX are recomputed by the next KKT step)
                double* lamX = C.dyX;
                K3R_LOOP for (int r = nZc + threadIdx.x; r < nFc; r += NT) C.tF[r] = C.wv[r];
                __syncthreads();
                tsolve(C, nZc, nFc, C.tF, lamG);
                ATmul(C, lamG, nullptr, C.t3);
                K3R_LOOP for (int k = threadIdx.x; k < nXc; k += NT) { const int j = C.xvar[k]; lamX[k] = (C.t4[j] - C.t3[j]) / C.xcv[k]; }
                __syncthreads();
                double lmax = 0;
                K3R_LOOP for (int r = nZc + threadIdx.x; r < nFc; r += NT) lmax = fmax(lmax, fabs(lamG[r]));
                K3R_LOOP for (int k = threadIdx.x; k < nXc; k += NT) lmax = fmax(lmax, fabs(lamX[k]));
                lmax = block_max(C, lmax);
                // dual ratio test of the exchange, two-pass; keyed by constraint index, packed with the position
                const double tdel = C.tie * C.ynoise;
                double thmax = INFINITY;
                K3R_LOOP for (int q = threadIdx.x; q < nXc + (nFc - nZc); q += NT) {
                    int c; double yc, lr;
                    if (q < nXc) { c = C.xcon[q]; yc = C.yX[q]; lr = lamX[q]; }
                    else { const int r = nFc - 1 - (q - nXc); c = C.widx[r]; yc = C.yW[r]; lr = lamG[r]; }
                    if (lb_new(C, c) == ub_new(C, c)) continue;
                    const double l = (side < 0) ? lr : -lr;
                    if (fabs(l) <= 1e-11 * (1.0 + lmax)) continue;
                    if (C.st[c] < 0 && l > 0) thmax = fmin(thmax, ((yc > 0 ? yc : 0) + tdel) * (1.0 / l));
                    else if (C.st[c] > 0 && l < 0) thmax = fmin(thmax, ((yc < 0 ? -yc : 0) + tdel) * (1.0 / -l));
                }
                thmax = block_min(C, thmax);
                int pk2 = RT_NONE;
                K3R_LOOP for (int q = threadIdx.x; q < nXc + (nFc - nZc); q += NT) {
                    int c; double yc, lr;
                    if (q < nXc) { c = C.xcon[q]; yc = C.yX[q]; lr = lamX[q]; }
                    else { const int r = nFc - 1 - (q - nXc); c = C.widx[r]; yc = C.yW[r]; lr = lamG[r]; }
                    if (lb_new(C, c) == ub_new(C, c)) continue;
                    const double l = (side < 0) ? lr : -lr;
                    if (fabs(l) <= 1e-11 * (1.0 + lmax)) continue;
                    if (C.st[c] < 0 && l > 0) { if ((yc > 0 ? yc : 0) * (1.0 / l) <= thmax) pk2 = min(pk2, c * 128 + q); }
                    else if (C.st[c] > 0 && l < 0) { if ((yc < 0 ? -yc : 0) * (1.0 / -l) <= thmax) pk2 = min(pk2, c * 128 + q); }
                }
                pk2 = block_imin(C, pk2);
                int kk = RT_NONE; double theta = INFINITY;
                if (pk2 != RT_NONE) {
                    kk = pk2 & 127;
                    double yc, lr;
                    if (kk < nXc) { yc = C.yX[kk]; lr = lamX[kk]; }
                    else { const int r = nFc - 1 - (kk - nXc); yc = C.yW[r]; lr = lamG[r]; }
                    const double l = (side < 0) ? lr : -lr;
                    const double num = (l > 0) ? (yc > 0 ? yc : 0) : (yc < 0 ? -yc : 0), den = fabs(l);
                    theta = (num <= tdel) ? 0.0 : num / den;
                    C.ynoise += theta * lmax;
                }
                if (kk == 0x7fffffff) { status = EEPACC_QP_INFEASIBLE; break; }
                K3R_LOOP for (int r = nZc + threadIdx.x; r < nFc; r += NT) C.yW[r] -= theta * ((side < 0) ? lamG[r] : -lamG[r]);
                K3R_LOOP for (int k = threadIdx.x; k < nXc; k += NT) C.yX[k] -= theta * ((side < 0) ? lamX[k] : -lamX[k]);
                __syncthreads();
                if (kk < nXc) {
                    if (threadIdx.x == 0 && zero_step) C.tabu[C.xcon[kk]] = 2;
                    __syncthreads();
                    if (free_var(C, kk)) { status = K3R_DEFER; break; }
                } else {
                    const int rp = nFc - 1 - (kk - nXc);
                    if (threadIdx.x == 0 && zero_step) C.tabu[C.widx[rp]] = 2;
                    __syncthreads();
                    remove_general(C, rp);
                }
                // new T row over the range space of the freed column; R untouched
                const int rn = C.nZ - 1;                 // freed column
                gatherF(C, C.t4);
                Qtmul(C, C.gF, C.tF, rn, C.nF);
                K3R_LOOP for (int c = rn + threadIdx.x; c < C.nF; c += NT) MM(rn, c) = C.tF[c];
                if (threadIdx.x == 0) {
                    C.widx[rn] = (short)blk; C.st[blk] = (signed char)side;
                    C.yW[rn] = (side < 0) ? theta : -theta;
                }
                C.nZ = rn;
                __syncthreads();
                n_exch++;
            }
            PH(15);
        } else {
            // ---------- dual blocking: remove constraint ----------
            const int q = qsel;
            int c, side;
            __syncthreads();
            if (q < nX) {
                c = C.xcon[q]; side = C.st[c];
                __syncthreads();
                if (threadIdx.x == 0 && zero_step) C.tabu[c] = 2;
                if (free_var(C, q)) { status = K3R_DEFER; break; }
            } else {
                const int rp = nF - 1 - (q - nX);
                c = C.widx[rp]; side = C.st[c];
                __syncthreads();
                if (threadIdx.x == 0 && zero_step) C.tabu[c] = 2;
                remove_general(C, rp);
            }
            int rc = resolve_pending(C, c, side, n_exch, zero_step);
            PH(16);
            if (rc) { status = rc; break; }
        }
    }

    if (status == K3R_DEFER) return K3R_DEFER;      // untouched warm state: the caller hands the vehicle over

    // ---------------- iterative refinement on the final working set ----------------
    if (status == EEPACC_QP_OK) {
        const int nZ = C.nZ, nF = C.nF, nX = C.nX;
        Amul(C, C.x, C.Adx);
        Hmul(C, C.x, C.t4);
        ATmul(C, C.yW, C.yX, C.t3);
        // rg = Hx + g - A_W' y ; rb = bound - a'x
        K3R_LOOP for (int j = threadIdx.x; j < n; j += NT) C.dg[j] = C.t4[j] + C.g[j] - C.t3[j];
        K3R_LOOP for (int r = nZ + threadIdx.x; r < nF; r += NT) {
            const int c = C.widx[r];
            C.dbw[r] = (C.st[c] < 0 ? C.lbc[c] : C.ubc[c]) - C.Adx[c];
        }
        K3R_LOOP for (int k = threadIdx.x; k < nX; k += NT) {
            const int c = C.xcon[k];
            C.dbX[k] = (C.st[c] < 0 ? C.lbc[c] : C.ubc[c]) - C.Adx[c];
        }
        __syncthreads();
        PH(17);
        kkt_step(C, C.dg, C.dbw, C.dbX, C.dx, C.dyW, C.dyX);
        K3R_LOOP for (int i = threadIdx.x; i < n; i += NT) C.x[i] += C.dx[i];
        K3R_LOOP for (int r = nZ + threadIdx.x; r < nF; r += NT) C.yW[r] += C.dyW[r];
        K3R_LOOP for (int k = threadIdx.x; k < nX; k += NT) C.yX[k] += C.dyX[k];
        __syncthreads();
    }
    PH(17);

    // ---------------- outputs + warm-start state ----------------
    {
        Hmul(C, C.x, C.t1);
        double s = 0;
        K3R_LOOP for (int i = threadIdx.x; i < n; i += NT) s += C.x[i] * (0.5 * (C.t1[i] - C.reg * C.x[i]) + C.g[i]);
        s = block_sum(C, s);
        const int nZ = C.nZ, nF = C.nF, nX = C.nX, nA = nX + nF - nZ;
        if (status == EEPACC_QP_OK) {
            int far = 0;
            K3R_LOOP for (int r = nZ + threadIdx.x; r < nF; r += NT) { int c = C.widx[r]; if (c >= m && fabs(C.x[c - m]) >= FARB) far = 1; }
            K3R_LOOP for (int k = threadIdx.x; k < nX; k += NT) { int c = C.xcon[k]; if (c >= m && fabs(C.x[c - m]) >= FARB) far = 1; }
            if (block_max(C, (double)far) > 0) status = EEPACC_QP_UNBOUNDED;
            int zr = 0;         // an all-zero row with inconsistent bounds: infeasible QP
            K3R_LOOP for (int c = threadIdx.x; c < m; c += NT)
                if (C.anorm[c] == 0.0 && (lb_new(C, c) > 1e-8 || ub_new(C, c) < -1e-8)) zr = 1;
            if (block_max(C, (double)zr) > 0) status = EEPACC_QP_INFEASIBLE;
        }
        // warm-start state in the layout of k3_qp.cu: position q lives at index n-1-q
        K3R_LOOP for (int i = threadIdx.x; i < n; i += NT) {
            io.x[veh * (size_t)n + i] = C.x[i];
            io.wx[veh * (size_t)n + i] = C.x[i];
            const int q = n - 1 - i;
            double yv = 0.0; short ci = -1; signed char sd = 0;
            if (q < nX) { yv = C.yX[q]; ci = C.xcon[q]; sd = C.st[ci]; }
            else if (q < nA) { const int r = nF - 1 - (q - nX); yv = C.yW[r]; ci = C.widx[r]; sd = C.st[ci]; }
            io.wy[veh * (size_t)n + i] = yv;
            io.widx[veh * (size_t)n + i] = ci;
            io.wside[veh * (size_t)n + i] = sd;
        }
        if (threadIdx.x == 0) {
            io.cost[veh] = s;
            io.status[veh] = (double)status;
            io.iter[veh] = (double)iter;
            io.wnA[veh] = nA;
            io.warm[veh] = (status == EEPACC_QP_OK) ? 1 : 0;
#ifdef EEP_K3_TIMERS
            PH(18);
            if (io.dbg && veh == 0) for (int q = 0; q < 28; ++q) io.dbg[(size_t)io.dbg_cap * 4 + q] = (double)C.ph[q];
#endif
            if (io.counters) {
                io.counters[veh * 4 + 0] += (unsigned long long)iter;
                io.counters[veh * 4 + 1] += (status != EEPACC_QP_OK);
                io.counters[veh * 4 + 2] += n_exch;
                io.counters[veh * 4 + 3] += n_refac;
            }
        }
    }
    return status;
}

template <int NU, int NB>
__device__ __forceinline__ void k3r_body(const eep_qp_io& io, unsigned char* k3r_smem)
{
    typedef k3r::CtxT<NU, NB> CT;
    CT C;
    k3r_setup<NU, NB>(C, io, k3r_smem, io.fc);
    k3r_tables(C, io);
    // hand-over launch (large capacity): only the vehicles the first level could not hold
    const long long n_work = io.todo_list ? (long long)*io.todo_cnt : io.B;
    for (long long wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
        const long long veh = io.todo_list ? (long long)io.todo_list[wi] : wi;
        if (k3r_vehicle(C, io, veh) == K3R_DEFER && threadIdx.x == 0) {
#ifdef EEP_HOST_EMU
            io.defer_list[(*io.defer_cnt)++] = (int)veh;
#else
            io.defer_list[atomicAdd(io.defer_cnt, 1)] = (int)veh;
#endif
        }
    }
}

template <int NU, int NB>
__global__ void __launch_bounds__(NT, K3R_MINB) k3r_qp_kernel(eep_qp_io io EMU_ONLY(, unsigned char* k3r_smem))
{
    k3r_body<NU, NB>(io, k3r_smem);
}

__global__ void k3r_rowtab_kernel(int ctrl, eepacc_params p, int m, unsigned long long* tab, double* consts, int* rc)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) *rc = k3r_build_rowtab(ctrl, p, m, tab, consts);
}

__host__ __device__ size_t eep_qpr_smem_base(int n, int m, int fc)
{
    const int mt = n + m, ldq = (fc + 1) | 1, NH = EEP_N;
    size_t d = (size_t)(fc + 1) * ldq + (size_t)fc * (fc + 1) / 2 + fc + 2 + (fc + 2) + 13 * (size_t)n + 10 * (size_t)(fc + 1)
             + (mt + 1) / 2
             + 2 * (size_t)(NH + 1) * NH + NH + 1 + 2 * NH * 9 + NH * 4 + 2 * (NH + 1) + NH + NH * 3
             + 2 * (NH + 1) + NH + 5 * (size_t)(fc + 1) + 16 + 4 + 96
             + (size_t)(2 * ((mt + NT - 1) / NT) + (n + NT - 1) / NT) * NT;
    size_t shorts = 7 * (size_t)(fc + 1) + 3 * (size_t)n;
    return d * 8 + shorts * 2 + 2 * (size_t)mt + 16;
}
extern "C" __host__ __device__ size_t eep_qpr_smem_bytes(int n, int m, int fc)
{
    size_t b = eep_qpr_smem_base(n, m, fc);
    if (eep_qpr_hs_in_smem(n, m, fc)) b += 32 * (size_t)(n + m);
#ifdef EEP_K3_TIMERS
    b = ((b + 7) & ~(size_t)7) + 34 * 8;
#endif
    return b;
}

#ifndef EEP_HOST_EMU
// one-off: packed row table of a controller into device memory; returns 0 or a CUDA / encoding error
extern "C" int eep_qpr_build_rowtab(int ctrl, const eepacc_params* p, int m, unsigned long long* tab_dev, double* consts_dev, cudaStream_t st)
{
    int* rc_dev = nullptr;
    cudaError_t e = cudaMalloc(&rc_dev, sizeof(int));
    if (e != cudaSuccess) return (int)e;
    k3r_rowtab_kernel<<<1, 32, 0, st>>>(ctrl, *p, m, tab_dev, consts_dev, rc_dev);
    int rc = -1;
    e = cudaMemcpyAsync(&rc, rc_dev, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(rc_dev);
    if (e != cudaSuccess) return (int)e;
    return rc ? 100000 + rc : 0;
}

template <int NU, int NB>
static int launch_qpr(const eep_qp_io* io, cudaStream_t st)
{
    const size_t sm = eep_qpr_smem_bytes(io->n, io->m, io->fc);
    static size_t optin[EEP_MAX_DEV];                    // per device and per instantiation (eep_optin_smem)
    if (int e = eep_optin_smem(k3r_qp_kernel<NU, NB>, sm, optin)) return e;
    long long g = io->B;               // one CTA per vehicle: the hardware scheduler balances uneven solves
    if (io->todo_list) {                         // hand-over list: a few persistent CTAs walk it
        const long long cap = io->todo_grid > 0 ? io->todo_grid : 148;
        if (g > cap) g = cap;
    }
    if (g < 1) g = 1;
    if (g > 2147483647LL) g = 2147483647LL;
    k3r_qp_kernel<NU, NB><<<(int)g, NT, sm, st>>>(*io);
    return (int)cudaGetLastError();
}

// =====================================================================================================
// K5: fused closed loop.  The per-step path (K2 launch -> K3r launch -> hand-over launches -> K4 launch)
// synchronises ALL vehicles of a slice at every kernel boundary, so every MPC step costs the slice its slowest
// vehicle (ncu, step 41 of a cycle: SMs 21 % active inside a K3r launch, the rest is the tail).  Here a CTA
// takes ONE vehicle through all its MPC steps; vehicles are handed out by an atomic counter; a solve that
// outgrows the shared-memory capacity is repeated in place at full capacity from a global work area.
// =====================================================================================================
#include "k2_build.cuh"
#include "k4_post.cuh"

template <int NU, int NB>
__global__ void __launch_bounds__(NT, K3R_MINB) k5_cycle_kernel(const eepacc_params p, const eep_cycle_io cy)
{
    typedef k3r::CtxT<NU, NB> CT;
    CT C;
    k3r_setup<NU, NB>(C, cy.q, k3r_smem, cy.q.fc);
    BuildSmem& S = *reinterpret_cast<BuildSmem*>(k3r_smem);        // K2's work area overlays the solver's
    __shared__ long long s_veh;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_veh = (long long)atomicAdd(cy.work_counter, 1ULL);
        __syncthreads();
        const long long veh = s_veh;
        if (veh >= cy.q.B) break;
        eep_build_io b = cy.b;
        eep_post_io po = cy.p;
        double t0 = cy.t0_begin;
        long long c_build = 0, c_qp = 0, c_big = 0, c_post = 0;     // thread 0: where this vehicle's cycles go
        for (int k = cy.k_begin; k < cy.k_end; ++k) {
            const long long tk0 = clock64();
            b.kk = k; b.t0 = t0;
            b.lead = cy.lead_all ? cy.lead_all + (size_t)k * cy.lead_stride : nullptr;
            k2_vehicle<NT>(p, b, veh, S);
            __syncthreads();                                       // K2's outputs (global) are visible to the CTA
            const long long tk1 = clock64();
            // ONE call site of the (force-inlined) solver: first in shared memory at capacity fc; a solve that outgrows it
            // is repeated in place at capacity n from this CTA's global work area (it cannot hand over again)
            long long tk2 = 0;
            for (int attempt = 0; attempt < 2; ++attempt) {
                if (attempt == 1) { k3r_setup<NU, NB>(C, cy.q, cy.big + (size_t)blockIdx.x * cy.big_bytes, cy.fc_big); __syncthreads(); }
                k3r_tables(C, cy.q);
                const int st = k3r_vehicle(C, cy.q, veh);
                if (attempt == 0) tk2 = clock64();
                if (st != K3R_DEFER) { if (attempt == 1) k3r_setup<NU, NB>(C, cy.q, k3r_smem, cy.q.fc); break; }
            }
            __syncthreads();
            const long long tk3 = clock64();
            if (threadIdx.x == 0) {
                po.kk = k;
                po.log = cy.log_all ? cy.log_all + (size_t)k * cy.log_stride : nullptr;
                po.lead = b.lead;
                k4_vehicle(p, po, veh);
                c_build += tk1 - tk0; c_qp += tk2 - tk1; c_big += tk3 - tk2; c_post += clock64() - tk3;
            }
            t0 += p.Ts;
        }
        if (threadIdx.x == 0 && cy.phase) {
            atomicAdd(cy.phase + 0, (unsigned long long)c_build); atomicAdd(cy.phase + 1, (unsigned long long)c_qp);
            atomicAdd(cy.phase + 2, (unsigned long long)c_big); atomicAdd(cy.phase + 3, (unsigned long long)c_post);
        }
    }
}

static size_t cycle_smem(const eep_cycle_io* cy)
{
    size_t sm = eep_qpr_smem_bytes(cy->q.n, cy->q.m, cy->q.fc);
    if (sm < sizeof(BuildSmem)) sm = sizeof(BuildSmem);
    return sm;
}
template <int NU, int NB>
static int cycle_prepare(const eep_cycle_io* cy, int* ctas_per_sm)
{
    const size_t sm = cycle_smem(cy);
    cudaError_t e = cudaFuncSetAttribute(k5_cycle_kernel<NU, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return (int)e;
    int nb = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k5_cycle_kernel<NU, NB>, NT, sm);
    if (e != cudaSuccess) return (int)e;
    *ctas_per_sm = nb;
    return 0;
}
extern "C" int eep_cycle_max_ctas_per_sm(const eep_cycle_io* cy)
{
    int nb = 0, rc;
    switch (cy->q.ctrl) {
    case EEPACC_FB: rc = cycle_prepare<6, 2>(cy, &nb); break;
    case EEPACC_AB: rc = cycle_prepare<5, 1>(cy, &nb); break;
    default: rc = cycle_prepare<2, 1>(cy, &nb); break;
    }
    return rc ? -1 : nb;
}
extern "C" int eep_launch_cycle(const eepacc_params* p, const eep_cycle_io* cy, int grid, cudaStream_t st)
{
    if (!cy->q.rowtab || !cy->q.rowc || !cy->q.sblk || !cy->q.hscr || !cy->work_counter || !cy->big) return (int)cudaErrorInvalidValue;
    const size_t sm = cycle_smem(cy);
    switch (cy->q.ctrl) {
    case EEPACC_FB: k5_cycle_kernel<6, 2><<<grid, NT, sm, st>>>(*p, *cy); break;
    case EEPACC_AB: k5_cycle_kernel<5, 1><<<grid, NT, sm, st>>>(*p, *cy); break;
    default: k5_cycle_kernel<2, 1><<<grid, NT, sm, st>>>(*p, *cy); break;
    }
    return (int)cudaGetLastError();
}

extern "C" int eep_launch_qpr(const eepacc_params*, const eep_qp_io* io, cudaStream_t st)
{
    if (!io->rowtab || !io->rowc || !io->sblk) return (int)cudaErrorInvalidValue;
    switch (io->ctrl) {
    case EEPACC_FB: return launch_qpr<6, 2>(io, st);
    case EEPACC_AB: return launch_qpr<5, 1>(io, st);
    default: return launch_qpr<2, 1>(io, st);
    }
}
#endif
