// k3_qp.cu -- K3: batched, warm-started, dense parametric active-set QP solver, one vehicle per CTA.
//
// Stands where the reference calls the CasADi conic plugin 'qpoases' (qpOASES 3.2 online
// active-set strategy; Ferreau et al., Math. Prog. Comp. 2014):
//     sol = QPsolver('h',H,'g',c,'a',G,'lbx',..,'ubx',..,'lba',g_lb,'uba',g_ub)
// (MC/RunOpt_ABMPC.m:252, RunOpt_FBMPC.m:278, RunOpt_BLMPC.m:230, RunOpt_TVMPC.m:208).
//
// Algorithm (same rules and tie-breaks as oracle/qp_oracle.c, which pins it against the golden runs):
//   * homotopy from the previous step's (x, y, working set) to the new data; primal/dual ratio
//     tests give the step length; one working-set change per stop;
//   * null-space factorisation  A_W Q = [0 T],  Z'HZ = R'R,  updated by Givens rotations;
//   * implicit far bounds +-1e6 on every variable and an all-fixed cold start keep Z'HZ positive
//     definite for semidefinite / indefinite / zero Hessians; a removal that would lose
//     definiteness becomes an exchange along the direction of non-positive curvature;
//   * linearly dependent additions are exchanged through a dual ratio test;
//   * H == 0 (the BL/TV LPs) is treated as 1e-6*I (least-norm LP optimum, cf. qpOASES HST_ZERO).
//
// Shared-memory plan (FP64, FB: n=120, m=522 -> 210 KB of the 227 KB per CTA):
//   Q      n x n column-major, leading dimension n|1 (odd: column walks are conflict-free)
//   M      packed upper triangle: rows/cols < nZ hold R, rows/cols >= nZ hold T
//          (T row of the constraint that owns Q column r is M row r), sub[] = one sub-diagonal
//   vectors x, g_cur, dx, ... (n) and Ax, lb_cur, ub_cur, A*dx (m+n)
// H, A, g and the bounds stay in HBM/L2: H*v and A*v are coalesced column-major sweeps, one
// row per thread.  Working-set position arrays are indexed by the Q column r they own.
#include "eepacc_dev.cuh"
#include "eepacc_kernels.h"
#include "eepacc_rows.cuh"

#ifndef K3_NT
#define K3_NT 256
#endif
#define NT K3_NT
#define FARB      1.0e6
#define RELAX     1.0e4
#define BIGB      1.0e19
#define EPS_LI    1.0e-10
#define EPS_PD    1.0e-9
#define EPS_RATE  1.0e-12
#define EPS_ROW   1.0e-12
#define FEAS_TOL  1.0e-10
#define LP_REG    1.0e-6

struct Ctx {
    int n, m, mt, ldq;
    int nA;
    // global
    const double *H, *g, *A, *lba, *uba, *lbx, *ubx;
    double reg, hscale;
    double tie, ynoise;     // ratio-test tie tolerance (0 for the LPs), round-off scale of the multipliers (qp_rules.cuh)
    // shared
    double *Q, *M, *sub;
    double *x, *gcur, *dx, *dg, *t1, *t2, *t3, *t4, *yW, *dyW, *dbw, *wv, *csA, *snA;
    double *Ax, *lbc, *ubc, *Adx;
    float* anorm;           // [mt] L1 norm of every row (only scales a noise threshold)
    // matrix-free mode (closed loop): K2's structured block staged in shared memory
    int smode, ctrl, nu, nb;
    const eepacc_params* p;
    double mass;
    double *PS, *PV, *vest, *Dc, *Cc, *Dg, *sv, *vv, *drv, *yc;
    short* widx;            // [n] by Q column r
    signed char* st;        // [mt]
    signed char* tabu;      // [mt] anti-cycling flags
    double* red;            // [2*NT/32 + 8]
    int* ired;
};

__device__ __forceinline__ int mi(const Ctx& c, int r, int col) { return r * c.n - (r * (r - 1)) / 2 + (col - r); }
#define MQ(i, j) C.Q[(size_t)(j) * C.ldq + (i)]
#define MM(r, c) C.M[mi(C, (r), (c))]

__device__ __forceinline__ double lb_new(const Ctx& C, int c)
{
    if (c < C.m) { double v = C.lba[c]; return (v <= -BIGB) ? -INFINITY : v; }
    double l = C.lbx ? C.lbx[c - C.m] : -INFINITY;
    return (l < -FARB) ? -FARB : l;
}
__device__ __forceinline__ double ub_new(const Ctx& C, int c)
{
    if (c < C.m) { double v = C.uba[c]; return (v >= BIGB) ? INFINITY : v; }
    double u = C.ubx ? C.ubx[c - C.m] : INFINITY;
    return (u > FARB) ? FARB : u;
}

// ---- block reductions (results broadcast to every thread) ----
__device__ __forceinline__ double block_sum(Ctx& C, double v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) C.red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0;
    for (int w = 0; w < NT / 32; ++w) s += C.red[w];
    return s;
}
__device__ __forceinline__ double block_max(Ctx& C, double v)
{
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) C.red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = C.red[0];
    for (int w = 1; w < NT / 32; ++w) s = fmax(s, C.red[w]);
    return s;
}
// lexicographic (t, key) minimum
__device__ __forceinline__ void block_argmin(Ctx& C, double& t, int& key)
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

#include "qp_rules.cuh"

struct VestSrc { const double* v_est; };

// ---- matrix-free operators (closed loop): H = Psi' Hs Psi and A = Gs Psi are never formed ----
// state sensitivities of a direction v: drv[i] = sum of the driven controls of stage i,
// sv[k] = d s_k, vv[k] = d v_k  (k = 0..N)
__device__ void s_prep(Ctx& C, const double* v)
{
    const int NH = EEP_N;
    for (int i = threadIdx.x; i < NH; i += NT) {
        double d = v[i * C.nu];
        if (C.nb == 2) d += v[i * C.nu + 1];
        C.drv[i] = d;
    }
    __syncthreads();
    for (int k = threadIdx.x; k <= NH; k += NT) {
        double a = 0, b = 0;
        for (int i = 0; i < k && i < NH; ++i) { a += C.PS[k * NH + i] * C.drv[i]; b += C.PV[k * NH + i] * C.drv[i]; }
        C.sv[k] = a; C.vv[k] = b;
    }
    __syncthreads();
}
__device__ __forceinline__ void s_rowdesc(const Ctx& C, int c, RowDesc& rd)
{
    VestSrc src{ C.vest };
    eep_row<false>(C.ctrl, *C.p, C.mass, src, c, rd);
}
// one entry A(c, j) of the dense constraint matrix
__device__ __forceinline__ double s_entry(const Ctx& C, const RowDesc& rd, int j)
{
    const int NH = EEP_N;
    const int i = j / C.nu, cj = j % C.nu;
    double val = 0.0;
    if (rd.kst >= 0 && cj < C.nb && i < rd.kst) val = rd.cS * C.PS[rd.kst * NH + i] + rd.cV * C.PV[rd.kst * NH + i];
    if (rd.cj0 == j) val += rd.cv0;
    if (rd.cj1 == j) val += rd.cv1;
    if (rd.cj2 == j) val += rd.cv2;
    if (rd.kst == -2 && cj < C.nb && i == rd.cj0 / C.nu - 1) val -= 1.0;
    return val;
}
__device__ __forceinline__ double A_entry(const Ctx& C, int c, int j)
{
    if (!C.smode) return C.A[c + (size_t)j * C.m];
    RowDesc rd;
    s_rowdesc(C, c, rd);
    return s_entry(C, rd, j);
}

// ---- dense primitives; every routine ends with data visible after a __syncthreads() ----
__device__ void get_row(Ctx& C, int c, double* a)
{
    const int n = C.n, m = C.m;
    if (c >= m) { for (int j = threadIdx.x; j < n; j += NT) a[j] = (j == c - m) ? 1.0 : 0.0; }
    else if (C.smode) {
        RowDesc rd;
        s_rowdesc(C, c, rd);
        for (int j = threadIdx.x; j < n; j += NT) a[j] = s_entry(C, rd, j);
    } else { for (int j = threadIdx.x; j < n; j += NT) a[j] = C.A[c + (size_t)j * m]; }
    __syncthreads();
}
__device__ void Hmul(Ctx& C, const double* v, double* out)
{
    const int n = C.n;
    if (C.smode) {
        const int NH = EEP_N, nu = C.nu, nb = C.nb;
        s_prep(C, v);
        // y_k on the core [v, driven]: D_k w_k + C_k w_{k-1} + C_{k+1}' w_{k+1}
        for (int k = threadIdx.x; k < NH; k += NT) {
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
        for (int j = threadIdx.x; j < n; j += NT) {
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
        return;
    }
    for (int i = threadIdx.x; i < n; i += NT) {
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        int j = 0;
        for (; j + 3 < n; j += 4) {
            s0 += C.H[i + (size_t)j * n] * v[j];
            s1 += C.H[i + (size_t)(j + 1) * n] * v[j + 1];
            s2 += C.H[i + (size_t)(j + 2) * n] * v[j + 2];
            s3 += C.H[i + (size_t)(j + 3) * n] * v[j + 3];
        }
        for (; j < n; ++j) s0 += C.H[i + (size_t)j * n] * v[j];
        out[i] = (s0 + s1) + (s2 + s3) + C.reg * v[i];
    }
    __syncthreads();
}
__device__ void Amul(Ctx& C, const double* v, double* out)
{
    const int n = C.n, m = C.m;
    if (C.smode) {
        s_prep(C, v);
        for (int r = threadIdx.x; r < m; r += NT) {
            RowDesc rd;
            s_rowdesc(C, r, rd);
            double val = 0.0;
            if (rd.kst >= 0) val = rd.cS * C.sv[rd.kst] + rd.cV * C.vv[rd.kst];
            if (rd.nctl > 0) val += rd.cv0 * v[rd.cj0];
            if (rd.nctl > 1) val += rd.cv1 * v[rd.cj1];
            if (rd.nctl > 2) val += rd.cv2 * v[rd.cj2];
            if (rd.kst == -2) val -= C.drv[rd.cj0 / C.nu - 1];
            out[r] = val;
        }
        for (int j = threadIdx.x; j < n; j += NT) out[m + j] = v[j];
        __syncthreads();
        return;
    }
    for (int r = threadIdx.x; r < m; r += NT) {
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        int j = 0;
        for (; j + 3 < n; j += 4) {
            s0 += C.A[r + (size_t)j * m] * v[j];
            s1 += C.A[r + (size_t)(j + 1) * m] * v[j + 1];
            s2 += C.A[r + (size_t)(j + 2) * m] * v[j + 2];
            s3 += C.A[r + (size_t)(j + 3) * m] * v[j + 3];
        }
        for (; j < n; ++j) s0 += C.A[r + (size_t)j * m] * v[j];
        out[r] = (s0 + s1) + (s2 + s3);
    }
    for (int j = threadIdx.x; j < n; j += NT) out[m + j] = v[j];
    __syncthreads();
}
// w[j] = q_j' a for j in [c0, c1)
__device__ void Qtmul(Ctx& C, const double* a, double* w, int c0, int c1)
{
    const int n = C.n;
    for (int j = c0 + threadIdx.x; j < c1; j += NT) {
        double s0 = 0, s1 = 0;
        int i = 0;
        for (; i + 1 < n; i += 2) { s0 += MQ(i, j) * a[i]; s1 += MQ(i + 1, j) * a[i + 1]; }
        if (i < n) s0 += MQ(i, j) * a[i];
        w[j] = s0 + s1;
    }
    __syncthreads();
}
// out[i] (+)= sum_{c in [c0,c1)} Q(i,c) p[c]
__device__ void Qmul(Ctx& C, const double* p, int c0, int c1, double* out, bool accumulate)
{
    const int n = C.n;
    for (int i = threadIdx.x; i < n; i += NT) {
        double s0 = 0, s1 = 0;
        int c = c0;
        for (; c + 1 < c1; c += 2) { s0 += MQ(i, c) * p[c]; s1 += MQ(i, c + 1) * p[c + 1]; }
        if (c < c1) s0 += MQ(i, c) * p[c];
        out[i] = (accumulate ? out[i] : 0.0) + (s0 + s1);
    }
    __syncthreads();
}
// solve M[lo:hi,lo:hi] x = b (back substitution); b is destroyed, x may not alias b.
// Blocks of 32 rows: one warp solves the diagonal block with shuffles (no block barrier inside),
// then all threads fold the solved block into the remaining right-hand side.
__device__ void usolve(Ctx& C, int lo, int hi, double* b, double* x)
{
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
        for (int i = lo + threadIdx.x; i < b0; i += NT) {
            double s0 = 0, s1 = 0;
            int c = b0;
            for (; c + 1 < b1; c += 2) { s0 += MM(i, c) * x[c]; s1 += MM(i, c + 1) * x[c + 1]; }
            if (c < b1) s0 += MM(i, c) * x[c];
            b[i] -= s0 + s1;
        }
    }
    __syncthreads();
}
// solve M[lo:hi,lo:hi]' y = v (forward substitution); v destroyed
__device__ void tsolve(Ctx& C, int lo, int hi, double* v, double* y)
{
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
        for (int i = b1 + threadIdx.x; i < hi; i += NT) {
            double s0 = 0, s1 = 0;
            int c = b0;
            for (; c + 1 < b1; c += 2) { s0 += MM(c, i) * y[c]; s1 += MM(c + 1, i) * y[c + 1]; }
            if (c < b1) s0 += MM(c, i) * y[c];
            v[i] -= s0 + s1;
        }
    }
    __syncthreads();
}

// Givens sweep that adds constraint row `a`.  nZt = number of null-space columns that take part
// (columns 0..nZt-1), nR = dimension of the valid R block to keep updated (0: none).  The new
// range-space column / T row is nZt-1.  Returns 1 if `a` is linearly dependent (no change).
__device__ int givens_add(Ctx& C, int cidx, int side, const double* a, int nZt, int nR, bool check_li)
{
    const int n = C.n;
    Qtmul(C, a, C.wv, 0, n);
    if (check_li) {
        double an = 0, zn = 0;
        for (int j = threadIdx.x; j < n; j += NT) { an += a[j] * a[j]; if (j < nZt) zn += C.wv[j] * C.wv[j]; }
        an = block_sum(C, an);
        zn = block_sum(C, zn);
        double san = sqrt(an);
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
    // Q: one row per thread, rotations applied left to right with the carry in a register
    for (int i = threadIdx.x; i < n; i += NT) {
        double u = MQ(i, 0);
        for (int k = 0; k < nZt - 1; ++k) {
            double v = MQ(i, k + 1);
            double cs = C.csA[k], sn = C.snA[k];
            MQ(i, k) = cs * u - sn * v;
            u = sn * u + cs * v;
        }
        if (nZt > 0) MQ(i, nZt - 1) = u;
    }
    if (nR > 1) {
        // R phase A: column rotations k = 0..nR-2, one row of R per thread
        for (int i = threadIdx.x; i < nR; i += NT) {
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
        // R phase B: re-triangularise (row rotations), wavefront over k.  (Measured: with up to 120 columns the whole
        // CTA per rotation beats the one-warp chain of qp_rules.cuh that the reduced-space kernel uses: 29.8 M vs
        // 32.1 M cycles of working-set changes per cold start.)
        for (int k = 0; k < nR - 1; ++k) {
            __syncthreads();
            double pa = MM(k, k), pb = C.sub[k + 1];
            double rr = hypot(pa, pb);
            __syncthreads();
            if (rr != 0.0) {
                double c2 = pa / rr, s2 = pb / rr;
                for (int j = k + threadIdx.x; j < nR; j += NT) {
                    double u = MM(k, j);
                    double v = (j == k) ? pb : MM(k + 1, j);
                    MM(k, j) = c2 * u + s2 * v;
                    if (j > k) MM(k + 1, j) = -s2 * u + c2 * v;
                }
            }
        }
    }
    __syncthreads();
    // new T row = M row nZt-1 (columns >= nZt-1)
    const int rn = nZt - 1;
    for (int c = rn + threadIdx.x; c < n; c += NT) MM(rn, c) = C.wv[c];
    if (threadIdx.x == 0) { C.widx[rn] = (short)cidx; C.st[cidx] = (signed char)side; }
    C.nA += 1;
    __syncthreads();
    return 0;
}

// Remove the constraint that owns Q column rp.  Afterwards column (n - nA_old) is the new
// null-space direction; R is not touched.  yW / widx are shifted with the rows.
__device__ void remove_constraint(Ctx& C, int rp)
{
    const int n = C.n, nZ = n - C.nA;
    if (threadIdx.x == 0) C.st[C.widx[rp]] = 0;
    __syncthreads();
    // shift rows nZ..rp-1 one row down the diagonal (row r-1 -> row r); one column per thread
    for (int c = nZ + threadIdx.x; c < n; c += NT) {
        if (c < rp) C.sub[c + 1] = MM(c, c);
        int rtop = c < rp ? c : rp;
        for (int r = rtop; r > nZ; --r) MM(r, c) = MM(r - 1, c);
    }
    if (threadIdx.x == 0) {
        for (int r = rp; r > nZ; --r) { C.widx[r] = C.widx[r - 1]; C.yW[r] = C.yW[r - 1]; }
    }
    // Givens wavefront: rotation at r (columns r-1, r) for r = rp .. nZ+1
    for (int r = rp; r > nZ; --r) {
        __syncthreads();
        double a = C.sub[r], b = MM(r, r);
        double rho = hypot(a, b);
        double cs = 1.0, sn = 0.0;
        if (rho != 0.0) { cs = b / rho; sn = a / rho; }
        __syncthreads();
        if (threadIdx.x == 0) { C.csA[r] = cs; C.snA[r] = sn; }
        // rows r' in [nZ+1, r]: elements (r', r-1) and (r', r)
        for (int rr = nZ + 1 + threadIdx.x; rr <= r; rr += NT) {
            double u = (rr == r) ? a : MM(rr, r - 1);
            double v = MM(rr, r);
            if (rr < r) MM(rr, r - 1) = cs * u - sn * v;
            MM(rr, r) = sn * u + cs * v;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += NT) {
        if (rp > nZ) {
            double v = MQ(i, rp);
            for (int r = rp; r > nZ; --r) {
                double u = MQ(i, r - 1);
                double cs = C.csA[r], sn = C.snA[r];
                MQ(i, r) = sn * u + cs * v;
                v = cs * u - sn * v;          // becomes column r-1, the right-hand column of the next rotation
            }
            MQ(i, nZ) = v;
        }
    }
    C.nA -= 1;
    __syncthreads();
}

// R = chol(Z'HZ) from scratch on the first nZ columns; returns 1 if not positive definite
__device__ int refactor_R(Ctx& C)
{
    const int n = C.n, nZ = n - C.nA;
    for (int j = 0; j < nZ; ++j) {
        Hmul(C, &MQ(0, j), C.t1);
        Qtmul(C, C.t1, C.t2, 0, j + 1);
        for (int i = threadIdx.x; i <= j; i += NT) MM(i, j) = C.t2[i];
    }
    const double thr = EPS_PD * C.hscale;
    for (int k = 0; k < nZ; ++k) {
        __syncthreads();
        double d = MM(k, k);
        if (!(d > thr)) return 1;
        double rk = sqrt(d), inv = 1.0 / rk;
        const int w = nZ - 1 - k;                 // trailing dimension
        for (int e = threadIdx.x; e < w * w; e += NT) {
            int i = k + 1 + e / w, j = k + 1 + e % w;
            if (j >= i) MM(i, j) -= (MM(k, i) * inv) * (MM(k, j) * inv);
        }
        __syncthreads();
        for (int j = k + threadIdx.x; j < nZ; j += NT) MM(k, j) = (j == k) ? rk : MM(k, j) * inv;
    }
    __syncthreads();
    return 0;
}

// KKT step: H dx + dg = A_W' dy,  A_W dx = db   (db, dy indexed by Q column)
__device__ void kkt_step(Ctx& C, const double* dg, const double* dbw, double* dx, double* dyW, long long* kp = nullptr)
{
    long long tk = clock64();
#define KPH(i) do { if (kp) { long long t_ = clock64(); kp[i] += t_ - tk; tk = t_; } } while (0)
    const int n = C.n, nZ = n - C.nA;
    double *p = C.t1, *u = C.t2, *t = C.t3;
    if (C.nA > 0) {
        for (int r = nZ + threadIdx.x; r < n; r += NT) t[r] = dbw[r];
        __syncthreads();
        usolve(C, nZ, n, t, p);
        KPH(0);
        Qmul(C, p, nZ, n, dx, false);
        KPH(1);
    } else {
        for (int i = threadIdx.x; i < n; i += NT) dx[i] = 0.0;
        __syncthreads();
    }
    Hmul(C, dx, u);
    for (int i = threadIdx.x; i < n; i += NT) u[i] += dg[i];
    __syncthreads();
    KPH(2);
    if (nZ > 0) {
        Qtmul(C, u, t, 0, nZ);
        for (int j = threadIdx.x; j < nZ; j += NT) t[j] = -t[j];
        __syncthreads();
        KPH(3);
        tsolve(C, 0, nZ, t, p);
        usolve(C, 0, nZ, p, t);
        KPH(4);
        Qmul(C, t, 0, nZ, p, false);            // p = Z q
        for (int i = threadIdx.x; i < n; i += NT) dx[i] += p[i];
        __syncthreads();
        KPH(5);
        Hmul(C, p, t);
        for (int i = threadIdx.x; i < n; i += NT) u[i] += t[i];
        __syncthreads();
        KPH(2);
    }
    if (C.nA > 0) {
        Qtmul(C, u, t, nZ, n);
        KPH(6);
        tsolve(C, nZ, n, t, dyW);
        KPH(7);
    }
}

// g_cur = A_W' y - H x  (optionally recomputing y by least squares first), duals sign-clamped
__device__ void reanchor(Ctx& C, bool recompute_y)
{
    const int n = C.n, m = C.m, nZ = n - C.nA;
    Hmul(C, C.x, C.t1);
    if (recompute_y && C.nA > 0) {
        for (int i = threadIdx.x; i < n; i += NT) C.t2[i] = C.t1[i] + C.gcur[i];
        __syncthreads();
        Qtmul(C, C.t2, C.t3, nZ, n);
        tsolve(C, nZ, n, C.t3, C.yW);
        double ym = C.ynoise;
        for (int r = nZ + threadIdx.x; r < n; r += NT) ym = fmax(ym, fabs(C.yW[r]));
        C.ynoise = block_max(C, ym);
    }
    for (int r = nZ + threadIdx.x; r < n; r += NT) {
        int c = C.widx[r];
        bool eq = lb_new(C, c) == ub_new(C, c);
        if (!eq) {
            if (C.st[c] < 0 && C.yW[r] < 0) C.yW[r] = 0;
            if (C.st[c] > 0 && C.yW[r] > 0) C.yW[r] = 0;
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += NT) {
        double s = -C.t1[j];
        for (int r = n - 1; r >= nZ; --r) {           // oldest first, like the oracle
            double yc = C.yW[r];
            if (yc == 0.0) continue;
            int c = C.widx[r];
            if (c < m) s += A_entry(C, c, j) * yc;
            else if (c - m == j) s += yc;
        }
        C.gcur[j] = s;
    }
    __syncthreads();
}

// see oracle/qp_oracle.c:resolve_pending
__device__ int resolve_pending(Ctx& C, int c_out, int side_out, unsigned long long& n_exch, bool zero_step)
{
    const int n = C.n, mt = C.mt;
    bool first = true, moved = false;
    for (;;) {
        const int kpd = n - C.nA - 1;
        if (kpd < 0) break;
        double* z = &MQ(0, kpd);
        double *hz = C.t1, *v = C.t2, *r = C.t3, *q = C.t4, *d = C.dx;
        Hmul(C, z, hz);
        double zhz = 0;
        for (int i = threadIdx.x; i < n; i += NT) zhz += z[i] * hz[i];
        zhz = block_sum(C, zhz);
        Qtmul(C, hz, v, 0, kpd);
        tsolve(C, 0, kpd, v, r);
        double rr = 0;
        for (int j = threadIdx.x; j < kpd; j += NT) rr += r[j] * r[j];
        rr = block_sum(C, rr);
        double rho2 = zhz - rr;
        if (rho2 > EPS_PD * C.hscale) {
            for (int j = threadIdx.x; j < kpd; j += NT) MM(j, kpd) = r[j];
            if (threadIdx.x == 0) MM(kpd, kpd) = sqrt(rho2);
            __syncthreads();
            break;
        }
        for (int j = threadIdx.x; j < kpd; j += NT) v[j] = r[j];
        __syncthreads();
        usolve(C, 0, kpd, v, q);
        Qmul(C, q, 0, kpd, d, false);
        for (int i = threadIdx.x; i < n; i += NT) d[i] = z[i] - d[i];
        __syncthreads();
        double sgn = 0;
        if (first) {
            get_row(C, c_out, C.t2);
            for (int i = threadIdx.x; i < n; i += NT) sgn += C.t2[i] * d[i];
            sgn = block_sum(C, sgn);
            if (side_out > 0) sgn = -sgn;
        } else {
            Hmul(C, C.x, hz);
            double gn = 0, dn2 = 0;
            for (int i = threadIdx.x; i < n; i += NT) { const double gi = hz[i] + C.gcur[i]; sgn -= gi * d[i]; gn += gi * gi; dn2 += d[i] * d[i]; }
            sgn = block_sum(C, sgn); gn = block_sum(C, gn); dn2 = block_sum(C, dn2);
            if (fabs(sgn) <= FLAT_TOL * sqrt(gn) * sqrt(dn2)) {
                // flat direction (gradient orthogonal to d within round-off): canonical sense, see qp_oracle.c
                get_row(C, c_out, C.t2);
                double an = 0;
                sgn = 0;
                for (int i = threadIdx.x; i < n; i += NT) { sgn += C.t2[i] * d[i]; an += C.t2[i] * C.t2[i]; }
                sgn = block_sum(C, sgn); an = block_sum(C, an);
                if (side_out > 0) sgn = -sgn;
                if (fabs(sgn) <= FLAT_TOL * sqrt(an) * sqrt(dn2)) {
                    double dmx = 0;
                    for (int i = threadIdx.x; i < n; i += NT) dmx = fmax(dmx, fabs(d[i]));
                    dmx = block_max(C, dmx);
                    int i0 = RT_NONE;
                    for (int i = threadIdx.x; i < n; i += NT) if (fabs(d[i]) > 0.25 * dmx) i0 = min(i0, i);
                    i0 = block_imin(C, i0);
                    sgn = (i0 == RT_NONE) ? 1.0 : d[i0];
                }
            }
        }
        if (sgn < 0) { __syncthreads(); for (int i = threadIdx.x; i < n; i += NT) d[i] = -d[i]; __syncthreads(); }
        Amul(C, d, C.Adx);
        double dn = 0, xmax = 0;
        for (int i = threadIdx.x; i < n; i += NT) { dn += d[i] * d[i]; xmax = fmax(xmax, fabs(C.x[i])); }
        dn = sqrt(block_sum(C, dn));
        xmax = block_max(C, xmax);
        const double pdel = C.tie * (1.0 + xmax);
        double tmax = INFINITY;
        for (int cc = threadIdx.x; cc < mt; cc += NT) {
            if (C.st[cc] != 0 || C.anorm[cc] == 0.0f) continue;
            int sd; double num, den;
            if (rt_exch(C, cc, dn, sd, num, den)) tmax = fmin(tmax, (num + pdel * C.anorm[cc]) * (1.0 / den));
        }
        tmax = block_min(C, tmax);
        if (!(tmax < INFINITY)) return EEPACC_QP_UNBOUNDED;
        int key = RT_NONE;
        for (int cc = threadIdx.x; cc < mt; cc += NT) {
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
        for (int i = threadIdx.x; i < n; i += NT) C.x[i] += tmin * d[i];
        for (int cc = threadIdx.x; cc < mt; cc += NT) C.Ax[cc] += tmin * C.Adx[cc];
        __syncthreads();
        if (threadIdx.x == 0) { if (jside < 0) C.lbc[j] = C.Ax[j]; else C.ubc[j] = C.Ax[j]; }
        get_row(C, j, C.t2);
        givens_add(C, j, jside, C.t2, kpd + 1, kpd, false);
        if (threadIdx.x == 0) { C.yW[kpd] = 0.0; if (zero_step) C.tabu[j] = 1; }
        __syncthreads();
        n_exch++;
        first = false; moved = true;
    }
    if (moved) reanchor(C, true);
    return 0;
}

// Warm start: rebuild  A_W Q = [0 T]  for the kept working set under the NEW constraint matrix.
//   A_W' = Qh [Rh; 0]  =>  Q = Qh with reversed columns,  T(i, c) = Rh(n-1-c, i).
// Most active rows are unit rows (slack / variable bounds, one entry): they are ordered first and
// eliminated by coordinate swaps (Qh = P_0 .. P_{nU-1} diag(I, G2)), so the Householder QR and the
// explicit formation of Qh only run on the few general rows in the remaining n - nU coordinates.
// Rows whose new bound is infinite, or that became linearly dependent, are dropped (y = 0).
__device__ void warm_setup(Ctx& C, int old_nA, const short* gidx, const signed char* gside, const double* gy, long long* wp = nullptr)
{
    long long tw = clock64();
#define WPH(i) do { if (wp) { long long t_ = clock64(); wp[i] += t_ - tw; tw = t_; } } while (0)
    const int n = C.n, mt = C.mt, m = C.m;
    for (int c = threadIdx.x; c < mt; c += NT) C.st[c] = 0;
    short* cand = reinterpret_cast<short*>(C.wv);        // [n] constraint index per QR column
    short* ucol = cand + n;                              // [n] unit rows: variable index (else -1)   (wv holds 4n shorts)
    short* swp = cand + 2 * n;                           // [n] swap partner of unit column k
    short* pos = cand + 3 * n;                           // [n] current row of coordinate j
    double* ycand = C.dyW;                               // [n]
    double* uval = C.t3;                                 // [n] coefficient of a unit row
    double* v0a = C.csA;                                 // v0 of every reflector
    double* beta = C.snA;                                // 2 / v'v
    double* rdiag = C.sub;                               // Rh(k,k)
    double* cn2 = C.dbw;                                 // squared norm of the original row
    short* tmpc = reinterpret_cast<short*>(C.t1);        // classification scratch: [3n] shorts
    double* tmpv = C.t2;
    __syncthreads();
    // 1. classify the stored working set in parallel (position k was stored at Q column n-1-k)
    for (int k = threadIdx.x; k < old_nA; k += NT) {
        const int r = n - 1 - k;
        const int c = gidx[r], sd = gside[r];
        const double bnd = sd < 0 ? lb_new(C, c) : ub_new(C, c);
        int jv = -1; double cv = 0.0; int ok = !isinf(bnd);
        if (ok) {
            if (c >= m) { jv = c - m; cv = 1.0; }
            else if (C.smode) {
                RowDesc rd; s_rowdesc(C, c, rd);
                if (rd.kst == -1 && rd.nctl == 1) { jv = rd.cj0; cv = rd.cv0; }
            } else {
                int nnz = 0, jl = -1; double vl = 0;
                for (int j = 0; j < n; ++j) { double v = C.A[c + (size_t)j * m]; if (v != 0.0) { ++nnz; jl = j; vl = v; } }
                if (nnz == 1) { jv = jl; cv = vl; }
            }
        }
        tmpc[3 * k] = (short)(ok ? c : -1); tmpc[3 * k + 1] = (short)jv; tmpc[3 * k + 2] = (short)sd;
        tmpv[k] = cv;
    }
    __syncthreads();
    WPH(0);
    // thread 0: ordered list (unit rows first, relative order kept) and the swap plan
    if (threadIdx.x == 0) {
        int cnt = 0;
        for (int j = 0; j < n; ++j) pos[j] = (short)j;
        short* inv = reinterpret_cast<short*>(C.t4);      // row -> coordinate
        for (int j = 0; j < n; ++j) inv[j] = (short)j;
        for (int k = 0; k < old_nA; ++k) {
            const int c = tmpc[3 * k], jv = tmpc[3 * k + 1];
            if (c < 0 || jv < 0) continue;
            const int a = pos[jv];
            if (a < cnt) continue;                        // coordinate already pinned: dependent, drop
            // swap rows a and cnt
            const int jo = inv[cnt];
            pos[jv] = (short)cnt; pos[jo] = (short)a; inv[cnt] = (short)jv; inv[a] = (short)jo;
            cand[cnt] = (short)c; ucol[cnt] = (short)jv; swp[cnt] = (short)a; uval[cnt] = tmpv[k];
            ycand[cnt] = gy[n - 1 - k]; C.st[c] = (signed char)tmpc[3 * k + 2];
            ++cnt;
        }
        C.ired[5] = cnt;                                  // nU
        for (int k = 0; k < old_nA; ++k) {
            const int c = tmpc[3 * k], jv = tmpc[3 * k + 1];
            if (c < 0 || jv >= 0) continue;
            cand[cnt] = (short)c; ucol[cnt] = -1; ycand[cnt] = gy[n - 1 - k]; C.st[c] = (signed char)tmpc[3 * k + 2];
            ++cnt;
        }
        C.ired[4] = cnt;
    }
    __syncthreads();
    int cnt = C.ired[4];
    const int nU = C.ired[5];
    WPH(1);
    // 2. stage the general rows as columns nU.. of W (= Q storage), apply the coordinate swaps
    for (int k = nU + threadIdx.x; k < cnt; k += NT) {
        const int c = cand[k];
        double s2 = 0;
        if (C.smode) {
            RowDesc rd;
            s_rowdesc(C, c, rd);
            for (int j = 0; j < n; ++j) { double v = s_entry(C, rd, j); MQ(j, k) = v; s2 += v * v; }
        } else {
            for (int j = 0; j < n; ++j) { double v = C.A[c + (size_t)j * m]; MQ(j, k) = v; s2 += v * v; }
        }
        cn2[k] = s2;
        for (int u = 0; u < nU; ++u) { const int a = swp[u]; if (a != u) { double t = MQ(u, k); MQ(u, k) = MQ(a, k); MQ(a, k) = t; } }
    }
    for (int k = threadIdx.x; k < nU; k += NT) rdiag[k] = uval[k];
    __syncthreads();
    WPH(2);
    // 3. Householder QR of the general columns on rows >= nU; a dependent column is swapped with the last and dropped
    int k = nU;
    while (k < cnt) {
        double s2 = 0;
        for (int i = k + threadIdx.x; i < n; i += NT) s2 += MQ(i, k) * MQ(i, k);
        s2 = block_sum(C, s2);
        const double ref = cn2[k] > 1.0 ? cn2[k] : 1.0;
        if (!(s2 > EPS_LI * EPS_LI * ref)) {
            const int last = cnt - 1;
            __syncthreads();
            if (threadIdx.x == 0) C.st[cand[k]] = 0;
            __syncthreads();
            if (k != last) {
                for (int i = threadIdx.x; i < n; i += NT) MQ(i, k) = MQ(i, last);
                if (threadIdx.x == 0) { cand[k] = cand[last]; ycand[k] = ycand[last]; cn2[k] = cn2[last]; }
            }
            cnt = last;
            __syncthreads();
            continue;
        }
        const double sigma = sqrt(s2), x0 = MQ(k, k);
        const double alpha = (x0 >= 0.0) ? -sigma : sigma;
        const double v0 = x0 - alpha;
        const double bt = 1.0 / (s2 - x0 * alpha);
        for (int j = k + 1 + threadIdx.x; j < cnt; j += NT) {
            double d = v0 * MQ(k, j), d1 = 0, d2 = 0, d3 = 0;
            int i = k + 1;
            for (; i + 3 < n; i += 4) {
                d += MQ(i, k) * MQ(i, j); d1 += MQ(i + 1, k) * MQ(i + 1, j);
                d2 += MQ(i + 2, k) * MQ(i + 2, j); d3 += MQ(i + 3, k) * MQ(i + 3, j);
            }
            for (; i < n; ++i) d += MQ(i, k) * MQ(i, j);
            d = ((d + d1) + (d2 + d3)) * bt;
            MQ(k, j) -= d * v0;
            for (int i2 = k + 1; i2 < n; ++i2) MQ(i2, j) -= d * MQ(i2, k);
        }
        if (threadIdx.x == 0) { v0a[k] = v0; beta[k] = bt; rdiag[k] = alpha; }
        __syncthreads();
        ++k;
    }
    WPH(3);
    const int nA = cnt, nZ = n - nA;
    // 4. T block of M from Rh:  M(r, c) = Rh(n-1-c, n-1-r),  c >= r >= nZ  (unit columns: diagonal only)
    for (int r = nZ + threadIdx.x; r < n; r += NT) {
        const int b = n - 1 - r;
        for (int c = r; c < n; ++c) {
            const int a = n - 1 - c;
            MM(r, c) = (a == b) ? rdiag[a] : (b < nU ? 0.0 : MQ(a, b));
        }
    }
    __syncthreads();
    WPH(4);
    // 5. form Qh = P_0 .. P_{nU-1} diag(I, G2): backward accumulation of the reflectors, then the swaps
    for (int j = threadIdx.x; j < n; j += NT)
        if (j < nU || j >= nA) { for (int i = 0; i < n; ++i) MQ(i, j) = (i == j) ? 1.0 : 0.0; }
    __syncthreads();
    for (int kk = nA - 1; kk >= nU; --kk) {
        const double v0 = v0a[kk], bt = beta[kk];
        for (int j = kk + 1 + threadIdx.x; j < n; j += NT) {
            double d = v0 * MQ(kk, j), d1 = 0, d2 = 0, d3 = 0;
            int i = kk + 1;
            for (; i + 3 < n; i += 4) {
                d += MQ(i, kk) * MQ(i, j); d1 += MQ(i + 1, kk) * MQ(i + 1, j);
                d2 += MQ(i + 2, kk) * MQ(i + 2, j); d3 += MQ(i + 3, kk) * MQ(i + 3, j);
            }
            for (; i < n; ++i) d += MQ(i, kk) * MQ(i, j);
            d = ((d + d1) + (d2 + d3)) * bt;
            MQ(kk, j) -= d * v0;
            for (int i2 = kk + 1; i2 < n; ++i2) MQ(i2, j) -= d * MQ(i2, kk);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += NT) {
            double vi = MQ(i, kk);
            MQ(i, kk) = (i < kk) ? 0.0 : (i == kk) ? 1.0 - bt * v0 * v0 : -bt * v0 * vi;
        }
        __syncthreads();
    }
    for (int j = threadIdx.x; j < n; j += NT)
        for (int u = nU - 1; u >= 0; --u) { const int a = swp[u]; if (a != u) { double t = MQ(u, j); MQ(u, j) = MQ(a, j); MQ(a, j) = t; } }
    __syncthreads();
    WPH(5);
    // 6. reverse the columns, 7. working-set bookkeeping by Q column
    for (int i = threadIdx.x; i < n; i += NT)
        for (int c = 0; c < n / 2; ++c) { double a = MQ(i, c), b = MQ(i, n - 1 - c); MQ(i, c) = b; MQ(i, n - 1 - c) = a; }
    for (int kq = threadIdx.x; kq < nA; kq += NT) { C.yW[n - 1 - kq] = ycand[kq]; C.widx[n - 1 - kq] = cand[kq]; }
    C.nA = nA;
    __syncthreads();
    WPH(6);
}

__device__ void cold_ws(Ctx& C)
{
    const int n = C.n, m = C.m, mt = C.mt;
    for (int c = threadIdx.x; c < mt; c += NT) C.st[c] = 0;
    for (int e = threadIdx.x; e < n * C.ldq; e += NT) C.Q[e] = 0.0;
    for (int e = threadIdx.x; e < n * (n + 1) / 2; e += NT) C.M[e] = 0.0;
    __syncthreads();
    for (int v = threadIdx.x; v < n; v += NT) {
        MQ(v, v) = 1.0;
        MM(v, v) = 1.0;                 // T row of variable v lives in M row v (Q column v)
        C.widx[v] = (short)(m + v);
        C.st[m + v] = -1;
        C.yW[v] = 0.5 + 0.5 * (double)(m + v) / (double)mt;   // distinct initial duals ("ramping")
    }
    C.nA = n;
    __syncthreads();
}

// auxiliary QP for which the current (x, y, working set) is optimal: relaxed bounds + anchored gradient
__device__ void setup_aux(Ctx& C)
{
    const int mt = C.mt;
    Amul(C, C.x, C.Ax);
    for (int c = threadIdx.x; c < mt; c += NT) {
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

extern __shared__ __align__(16) unsigned char k3_smem[];

__global__ void __launch_bounds__(NT, 1) k3_qp_kernel(const __grid_constant__ eepacc_params prm, eep_qp_io io)
{
    Ctx C;
    const int n = io.n, m = io.m, mt = n + m;
    C.n = n; C.m = m; C.mt = mt; C.ldq = n | 1;
    {
        double* p = reinterpret_cast<double*>(k3_smem);
        C.Q = p; p += (size_t)n * C.ldq;
        C.M = p; p += (size_t)n * (n + 1) / 2;
        C.sub = p; p += n + 1;
        C.x = p; p += n; C.gcur = p; p += n; C.dx = p; p += n; C.dg = p; p += n;
        C.t1 = p; p += n; C.t2 = p; p += n; C.t3 = p; p += n; C.t4 = p; p += n;
        C.yW = p; p += n; C.dyW = p; p += n; C.dbw = p; p += n; C.wv = p; p += n;
        C.csA = p; p += n + 1; C.snA = p; p += n + 1;
        C.Ax = p; p += mt; C.lbc = p; p += mt; C.ubc = p; p += mt; C.Adx = p; p += mt;
        C.anorm = reinterpret_cast<float*>(p); p += (mt + 1) / 2;
        C.smode = io.sblk != nullptr; C.ctrl = io.ctrl; C.p = &prm;
        C.nu = n / EEP_N; C.nb = (io.ctrl == EEPACC_FB) ? 2 : 1;
        if (C.smode) {
            const int NH = EEP_N;
            C.PS = p; p += (NH + 1) * NH; C.PV = p; p += (NH + 1) * NH; C.vest = p; p += NH;
            C.Dc = p; p += NH * 9; C.Cc = p; p += NH * 9; C.Dg = p; p += NH * 4;
            C.sv = p; p += NH + 1; C.vv = p; p += NH + 1; C.drv = p; p += NH; C.yc = p; p += NH * 3;
        }
        C.red = p; p += 8;
        C.ired = reinterpret_cast<int*>(p); p += 4;      // 8 ints
        C.widx = reinterpret_cast<short*>(p); p += (n + 3) / 4 + 1;
        C.st = reinterpret_cast<signed char*>(p);
        C.tabu = C.st + mt;
    }
    // fallback mode (closed loop): only the vehicles the reduced-space kernel (k3r_qp.cu) handed over
    const long long n_work = io.todo_list ? (long long)*io.todo_cnt : io.B;
    for (long long wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
        const long long veh = io.todo_list ? (long long)io.todo_list[wi] : wi;
        __syncthreads();
        C.H = io.H ? io.H + veh * (size_t)n * n : nullptr;
        C.g = io.g + veh * (size_t)n;
        C.A = io.A ? io.A + veh * (size_t)m * n : nullptr;
        C.lba = io.lba + veh * (size_t)m;
        C.uba = io.uba + veh * (size_t)m;
        C.lbx = io.lbx ? io.lbx + veh * (size_t)n : nullptr;
        C.ubx = io.ubx ? io.ubx + veh * (size_t)n : nullptr;
        int status = EEPACC_QP_OK, iter = 0;
        const int max_iter = 5 * (n + m);
        const int restart_after = (4 * n > 400 ? 4 * n : 400);
#ifdef EEP_K3_TIMERS
        long long ph[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }, tck = clock64();
        long long kph[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }, rph[4] = { 0, 0, 0, 0 }, rph2[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
#define PHASE(i) do { long long t_ = clock64(); ph[i] += t_ - tck; tck = t_; } while (0)
#define KPHP kph
#define WPHP rph2
#define RPH(i) do { rph[i] += clock64() - tr; tr = clock64(); } while (0)
#else
#define PHASE(i) do { } while (0)
#define KPHP nullptr
#define WPHP nullptr
#define RPH(i) do { } while (0)
#endif
        unsigned long long n_exch = 0, n_refac = 0;

        // Hessian scale, LP regularisation, row norms
        if (C.smode) {
            const double* sb = io.sblk + veh * (size_t)EEP_SBLK;
            for (int e = threadIdx.x; e < (EEP_N + 1) * EEP_N; e += NT) { C.PS[e] = sb[EEP_SB_PS + e]; C.PV[e] = sb[EEP_SB_PV + e]; }
            for (int e = threadIdx.x; e < EEP_N * 9; e += NT) { C.Dc[e] = sb[EEP_SB_DC + e]; C.Cc[e] = sb[EEP_SB_CC + e]; }
            for (int e = threadIdx.x; e < EEP_N * 4; e += NT) C.Dg[e] = sb[EEP_SB_DG + e];
            for (int e = threadIdx.x; e < EEP_N; e += NT) C.vest[e] = sb[EEP_SB_VEST + e];
            C.reg = (sb[EEP_SB_HMAX] == 0.0) ? LP_REG : 0.0;
            C.hscale = fmax(sb[EEP_SB_HSCALE], LP_REG);
            C.tie = (C.reg != 0.0) ? 0.0 : TIE_TOL;
            C.mass = sb[EEP_SB_MASS];
            __syncthreads();
            for (int r = threadIdx.x; r < m; r += NT) {
                RowDesc rd;
                s_rowdesc(C, r, rd);
                double sa = 0;
                if (rd.kst >= 0)
                    for (int i = 0; i < rd.kst && i < EEP_N; ++i)
                        sa += C.nb * fabs(rd.cS * C.PS[rd.kst * EEP_N + i] + rd.cV * C.PV[rd.kst * EEP_N + i]);
                sa += fabs(rd.cv0) + fabs(rd.cv1) + fabs(rd.cv2);
                if (rd.kst == -2) sa += C.nb;
                C.anorm[r] = (float)sa;
            }
        } else {
            double hd = 0, hm = 0;
            for (int e = threadIdx.x; e < n * n; e += NT) { double v = fabs(C.H[e]); hm = fmax(hm, v); if (e % (n + 1) == 0) hd = fmax(hd, v); }
            hd = block_max(C, hd);
            hm = block_max(C, hm);
            C.reg = (hm == 0.0) ? LP_REG : 0.0;
            C.hscale = fmax(hd, LP_REG);
            C.tie = (C.reg != 0.0) ? 0.0 : TIE_TOL;
            for (int r = threadIdx.x; r < m; r += NT) {
                double sa = 0;
                for (int j = 0; j < n; ++j) sa += fabs(C.A[r + (size_t)j * m]);
                C.anorm[r] = (sa > 0.0 && sa < 1e-30) ? 1e-30f : (float)sa;
            }
        }
        for (int j = threadIdx.x; j < n; j += NT) C.anorm[m + j] = 1.0f;
        for (int c = threadIdx.x; c < mt; c += NT) C.tabu[c] = 0;
        __syncthreads();

        PHASE(0);
        // ---------------- set-up ----------------
        const int was_warm = io.warm[veh];
        bool need_cold = !was_warm;
        for (int i = threadIdx.x; i < n; i += NT) C.x[i] = was_warm ? io.wx[veh * (size_t)n + i] : 0.0;
        if (!need_cold) {
            warm_setup(C, io.wnA[veh], io.widx + veh * (size_t)n, io.wside + veh * (size_t)n, io.wy + veh * (size_t)n, WPHP);
            PHASE(1);
            n_refac++;
            if (refactor_R(C)) need_cold = true;
            PHASE(2);
        }
        if (need_cold) cold_ws(C);
        setup_aux(C);
        {
            double ym = 1.0;
            for (int r = n - C.nA + threadIdx.x; r < n; r += NT) ym = fmax(ym, fabs(C.yW[r]));
            C.ynoise = block_max(C, ym);
        }
        PHASE(3);

        // ---------------- homotopy ----------------
        for (;;) {
            const int nZ = n - C.nA;
            for (int i = threadIdx.x; i < n; i += NT) C.dg[i] = C.g[i] - C.gcur[i];
            for (int r = nZ + threadIdx.x; r < n; r += NT) {
                int c = C.widx[r];
                double nb = C.st[c] < 0 ? lb_new(C, c) : ub_new(C, c);
                double cb = C.st[c] < 0 ? C.lbc[c] : C.ubc[c];
                C.dbw[r] = isinf(nb) ? 0.0 : nb - cb;
            }
            __syncthreads();
            kkt_step(C, C.dg, C.dbw, C.dx, C.dyW, KPHP);
            PHASE(4);
#ifdef EEP_K3_TIMERS
            long long tr = clock64();
#endif
            Amul(C, C.dx, C.Adx);
            RPH(0);

            // ---- ratio tests, two-pass (qp_rules.cuh; oracle/qp_oracle.c:rt_select) ----
            double dxmax = 0, xmax = 0;
            for (int i = threadIdx.x; i < n; i += NT) { dxmax = fmax(dxmax, fabs(C.dx[i])); xmax = fmax(xmax, fabs(C.x[i])); }
            block_max2(C, dxmax, xmax);
            double dys = 0;
            for (int r = nZ + threadIdx.x; r < n; r += NT) dys = fmax(dys, fabs(C.dyW[r]));
            dys = block_max(C, dys);
            const double thr = EPS_RATE * (1.0 + dys);
            const double pdel = C.tie * (1.0 + xmax), ddel = C.tie * C.ynoise;
            double tmax = INFINITY;
            for (int c = threadIdx.x; c < mt; c += NT) {
                if (C.st[c] != 0 || C.anorm[c] == 0.0f) continue;      // all-zero rows can never be added
                if (C.tabu[c] == 2) continue;                          // left during this run of zero steps: no re-entry
                double num, den;
                if (rt_primal(C, c, 0, dxmax, num, den)) tmax = fmin(tmax, (num + pdel * C.anorm[c]) * (1.0 / den));
                if (rt_primal(C, c, 1, dxmax, num, den)) tmax = fmin(tmax, (num + pdel * C.anorm[c]) * (1.0 / den));
            }
            for (int r = nZ + threadIdx.x; r < n; r += NT) {
                const int c = C.widx[r];
                if (lb_new(C, c) == ub_new(C, c)) continue;
                if (C.tabu[c] == 1) continue;            // entered during this run of zero steps: stays
                double num, den;
                if (rt_dual(C.st[c], C.yW[r], C.dyW[r], thr, ddel, num, den)) tmax = fmin(tmax, (num + ddel) * (1.0 / den));
            }
            tmax = block_min(C, tmax);
            // pass 2: lowest key among the candidates within tmax; packed = key * 128 + Q column of a dual candidate
            int pk = RT_NONE;
            if (tmax < INFINITY) {
                for (int c = threadIdx.x; c < mt; c += NT) {
                    if (C.st[c] != 0 || C.anorm[c] == 0.0f) continue;
                    if (C.tabu[c] == 2) continue;
                    double num, den;
                    if (rt_primal(C, c, 0, dxmax, num, den) && num * (1.0 / den) <= tmax) pk = min(pk, (2 * c) * 128);
                    if (rt_primal(C, c, 1, dxmax, num, den) && num * (1.0 / den) <= tmax) pk = min(pk, (2 * c + 1) * 128);
                }
                for (int r = nZ + threadIdx.x; r < n; r += NT) {
                    const int c = C.widx[r];
                    if (lb_new(C, c) == ub_new(C, c)) continue;
                    if (C.tabu[c] == 1) continue;
                    double num, den;
                    if (rt_dual(C.st[c], C.yW[r], C.dyW[r], thr, ddel, num, den) && num * (1.0 / den) <= tmax) pk = min(pk, (2 * mt + c) * 128 + r);
                }
                pk = block_imin(C, pk);
            }
            RPH(1);
            const bool full = (pk == RT_NONE);
            double tau = 1.0;
            const int key = full ? RT_NONE : pk >> 7, rsel = pk & 127;
            if (!full) {
                double num = 0, den = 1, del;
                if (key < 2 * mt) { rt_primal(C, key >> 1, key & 1, dxmax, num, den); del = pdel * C.anorm[key >> 1]; }
                else { rt_dual(C.st[C.widx[rsel]], C.yW[rsel], C.dyW[rsel], thr, ddel, num, den); del = ddel; }
                tau = (num <= del) ? 0.0 : num / den;       // within round-off of zero: a zero step
            }
            C.ynoise += tau * dys;
            // anti-cycling (see oracle/qp_oracle.c): constraints that enter while the homotopy makes no
            // progress may not leave through the dual test until it does
            const bool zero_step = !(tau > 1e-14);
            if (!zero_step) for (int c = threadIdx.x; c < mt; c += NT) C.tabu[c] = 0;

            // take the step
            for (int i = threadIdx.x; i < n; i += NT) {
                C.x[i] += tau * C.dx[i];
                C.gcur[i] = full ? C.g[i] : C.gcur[i] + tau * C.dg[i];
            }
            for (int r = nZ + threadIdx.x; r < n; r += NT) C.yW[r] += tau * C.dyW[r];
            for (int c = threadIdx.x; c < mt; c += NT) {
                C.Ax[c] += tau * C.Adx[c];
                double ln = lb_new(C, c), un = ub_new(C, c);
                if (!isinf(ln)) C.lbc[c] = full ? ln : C.lbc[c] + tau * (ln - C.lbc[c]);
                if (!isinf(un)) C.ubc[c] = full ? un : C.ubc[c] + tau * (un - C.ubc[c]);
            }
            __syncthreads();
            RPH(2);
            PHASE(5);
            if (full) break;
            if (++iter > max_iter) { status = EEPACC_QP_MAXITER; break; }
            if (iter == restart_after) {
                // the warm-started path wanders (degenerate / non-convex stretches): start over from an
                // all-fixed working set at the current point -- the cold-start path
                cold_ws(C);
                setup_aux(C);
                for (int c = threadIdx.x; c < mt; c += NT) C.tabu[c] = 0;
                __syncthreads();
                continue;
            }
            if (io.dbg && veh == 0 && threadIdx.x == 0 && iter <= io.dbg_cap) {
                double* o = io.dbg + (size_t)(iter - 1) * 4;
                o[0] = tau; o[1] = (double)key; o[2] = (double)C.nA; o[3] = (double)n_exch;
            }

            if (key < 2 * mt) {
                // ---------- primal blocking: add constraint ----------
                const int blk = key >> 1, side = (key & 1) ? 1 : -1;
                if (threadIdx.x == 0) {
                    if (side < 0) C.lbc[blk] = C.Ax[blk]; else C.ubc[blk] = C.Ax[blk];
                    if (zero_step) C.tabu[blk] = 1;
                }
                get_row(C, blk, C.t4);
                const int nZc = n - C.nA;
                if (givens_add(C, blk, side, C.t4, nZc, nZc, true) == 0) {
                    if (threadIdx.x == 0) C.yW[nZc - 1] = 0.0;
                    __syncthreads();
                } else {
                    // linearly dependent: a = A_W' lam, exchange by a dual ratio test
                    Qtmul(C, C.t4, C.t1, nZc, n);
                    tsolve(C, nZc, n, C.t1, C.t2);          // lam, by Q column
                    double lmax = 0;
                    for (int r = nZc + threadIdx.x; r < n; r += NT) lmax = fmax(lmax, fabs(C.t2[r]));
                    lmax = block_max(C, lmax);
                    // dual ratio test of the exchange, two-pass; keyed by constraint index, packed with the Q column
                    const double tdel = C.tie * C.ynoise;
                    double thmax = INFINITY;
                    for (int r = nZc + threadIdx.x; r < n; r += NT) {
                        const int c = C.widx[r];
                        if (lb_new(C, c) == ub_new(C, c)) continue;
                        const double l = (side < 0) ? C.t2[r] : -C.t2[r];
                        if (fabs(l) <= 1e-11 * (1.0 + lmax)) continue;
                        const double yc = C.yW[r];
                        if (C.st[c] < 0 && l > 0) thmax = fmin(thmax, ((yc > 0 ? yc : 0) + tdel) * (1.0 / l));
                        else if (C.st[c] > 0 && l < 0) thmax = fmin(thmax, ((yc < 0 ? -yc : 0) + tdel) * (1.0 / -l));
                    }
                    thmax = block_min(C, thmax);
                    int pk2 = RT_NONE;
                    for (int r = nZc + threadIdx.x; r < n; r += NT) {
                        const int c = C.widx[r];
                        if (lb_new(C, c) == ub_new(C, c)) continue;
                        const double l = (side < 0) ? C.t2[r] : -C.t2[r];
                        if (fabs(l) <= 1e-11 * (1.0 + lmax)) continue;
                        const double yc = C.yW[r];
                        if (C.st[c] < 0 && l > 0) { if ((yc > 0 ? yc : 0) * (1.0 / l) <= thmax) pk2 = min(pk2, c * 128 + r); }
                        else if (C.st[c] > 0 && l < 0) { if ((yc < 0 ? -yc : 0) * (1.0 / -l) <= thmax) pk2 = min(pk2, c * 128 + r); }
                    }
                    pk2 = block_imin(C, pk2);
                    int kk = RT_NONE; double theta = INFINITY;
                    if (pk2 != RT_NONE) {
                        const int r = pk2 & 127;
                        kk = n - 1 - r;
                        const double yc = C.yW[r], l = (side < 0) ? C.t2[r] : -C.t2[r];
                        const double num = (l > 0) ? (yc > 0 ? yc : 0) : (yc < 0 ? -yc : 0), den = fabs(l);
                        theta = (num <= tdel) ? 0.0 : num / den;
                        C.ynoise += theta * lmax;
                    }
                    if (kk == 0x7fffffff) { status = EEPACC_QP_INFEASIBLE; break; }
                    for (int r = nZc + threadIdx.x; r < n; r += NT) {
                        double l = (side < 0) ? C.t2[r] : -C.t2[r];
                        C.yW[r] -= theta * l;
                    }
                    if (threadIdx.x == 0 && zero_step) C.tabu[C.widx[n - 1 - kk]] = 2;
                    __syncthreads();
                    remove_constraint(C, n - 1 - kk);
                    // new T row over the (unchanged-size) range space; R untouched
                    const int rn = nZc;                      // freed column
                    Qtmul(C, C.t4, C.t1, rn, n);
                    for (int c = rn + threadIdx.x; c < n; c += NT) MM(rn, c) = C.t1[c];
                    if (threadIdx.x == 0) {
                        C.widx[rn] = (short)blk; C.st[blk] = (signed char)side;
                        C.yW[rn] = (side < 0) ? theta : -theta;
                    }
                    C.nA += 1;
                    __syncthreads();
                    n_exch++;
                }
            } else {
                // ---------- dual blocking: remove constraint ----------
                const int rp = rsel;
                const int c = C.widx[rp];
                const int side = C.st[c];
                __syncthreads();
                if (threadIdx.x == 0 && zero_step) C.tabu[c] = 2;
                remove_constraint(C, rp);
                int rc = resolve_pending(C, c, side, n_exch, zero_step);
                if (rc) { status = rc; break; }
            }
            PHASE(6);
        }

        PHASE(6);
        // ---------------- iterative refinement on the final working set ----------------
        if (status == EEPACC_QP_OK) {
            for (int pass = 0; pass < 1; ++pass) {
                const int nZ = n - C.nA;
                Amul(C, C.x, C.Adx);
                Hmul(C, C.x, C.t4);
                // rg = Hx + g - A_W' y ; rb = bound - a'x
                for (int j = threadIdx.x; j < n; j += NT) {
                    double s = C.t4[j] + C.g[j];
                    for (int r = n - 1; r >= nZ; --r) {
                        double yc = C.yW[r];
                        int c = C.widx[r];
                        if (c < m) s -= A_entry(C, c, j) * yc;
                        else if (c - m == j) s -= yc;
                    }
                    C.dg[j] = s;
                }
                for (int r = nZ + threadIdx.x; r < n; r += NT) {
                    const int c = C.widx[r];
                    C.dbw[r] = (C.st[c] < 0 ? C.lbc[c] : C.ubc[c]) - C.Adx[c];
                }
                __syncthreads();
                kkt_step(C, C.dg, C.dbw, C.dx, C.dyW);
                for (int i = threadIdx.x; i < n; i += NT) C.x[i] += C.dx[i];
                for (int r = nZ + threadIdx.x; r < n; r += NT) C.yW[r] += C.dyW[r];
                __syncthreads();
            }
        }

        // ---------------- outputs + warm-start state ----------------
        {
            Hmul(C, C.x, C.t1);
            double s = 0;
            for (int i = threadIdx.x; i < n; i += NT) s += C.x[i] * (0.5 * (C.t1[i] - C.reg * C.x[i]) + C.g[i]);
            s = block_sum(C, s);
            const int nZ = n - C.nA;
            if (status == EEPACC_QP_OK) {
                int far = 0;
                for (int r = nZ + threadIdx.x; r < n; r += NT) { int c = C.widx[r]; if (c >= m && fabs(C.x[c - m]) >= FARB) far = 1; }
                if (block_max(C, (double)far) > 0) status = EEPACC_QP_UNBOUNDED;
                int zr = 0;         // an all-zero row with inconsistent bounds: infeasible QP
                for (int c = threadIdx.x; c < m; c += NT)
                    if (C.anorm[c] == 0.0 && (lb_new(C, c) > 1e-8 || ub_new(C, c) < -1e-8)) zr = 1;
                if (block_max(C, (double)zr) > 0) status = EEPACC_QP_INFEASIBLE;
            }
            for (int i = threadIdx.x; i < n; i += NT) {
                io.x[veh * (size_t)n + i] = C.x[i];
                io.wx[veh * (size_t)n + i] = C.x[i];
                if (io.lam_x) io.lam_x[veh * (size_t)n + i] = 0.0;
                io.wy[veh * (size_t)n + i] = (i >= nZ) ? C.yW[i] : 0.0;
                io.widx[veh * (size_t)n + i] = (i >= nZ) ? C.widx[i] : (short)-1;
                io.wside[veh * (size_t)n + i] = (i >= nZ) ? C.st[C.widx[i]] : (signed char)0;
            }
            if (io.lam_a) for (int c = threadIdx.x; c < m; c += NT) io.lam_a[veh * (size_t)m + c] = 0.0;
            __syncthreads();
            // CasADi sign convention (multiplier > 0 at an upper bound) = -qpOASES'
            for (int r = nZ + threadIdx.x; r < n; r += NT) {
                int c = C.widx[r];
                if (c < m) { if (io.lam_a) io.lam_a[veh * (size_t)m + c] = -C.yW[r]; }
                else if (io.lam_x) io.lam_x[veh * (size_t)n + (c - m)] = -C.yW[r];
            }
            if (threadIdx.x == 0) {
                io.cost[veh] = s;
                io.status[veh] = (double)status;
                io.iter[veh] = (double)iter;
                io.wnA[veh] = C.nA;
                io.warm[veh] = (status == EEPACC_QP_OK) ? 1 : 0;
                PHASE(7);
#ifdef EEP_K3_TIMERS
                if (io.dbg && veh == 0) {
                    for (int q = 0; q < 8; ++q) io.dbg[(size_t)io.dbg_cap * 4 + q] = (double)ph[q];
                    for (int q = 0; q < 8; ++q) io.dbg[(size_t)io.dbg_cap * 4 + 8 + q] = (double)kph[q];
                    for (int q = 0; q < 4; ++q) io.dbg[(size_t)io.dbg_cap * 4 + 16 + q] = (double)rph[q];
                    for (int q = 0; q < 8; ++q) io.dbg[(size_t)io.dbg_cap * 4 + 20 + q] = (double)rph2[q];
                }
#endif
                if (io.counters) {
                    io.counters[veh * 4 + 0] += (unsigned long long)iter;
                    io.counters[veh * 4 + 1] += (status != EEPACC_QP_OK);
                    io.counters[veh * 4 + 2] += n_exch;
                    io.counters[veh * 4 + 3] += n_refac;
                }
            }
        }
    }
}

extern "C" size_t eep_qp_smem_bytes(int n, int m, int smode)
{
    const int mt = n + m, ldq = n | 1, NH = EEP_N;
    size_t d = (size_t)n * ldq + (size_t)n * (n + 1) / 2 + (n + 1) + 12 * (size_t)n + 2 * (size_t)(n + 1)
             + 4 * (size_t)mt + (mt + 1) / 2 + 8 + 4 + ((n + 3) / 4 + 1);
    if (smode) d += 2 * (size_t)(NH + 1) * NH + NH + 2 * NH * 9 + NH * 4 + 2 * (NH + 1) + NH + NH * 3;
    return d * 8 + 2 * (size_t)mt + 16;
}

extern "C" int eep_launch_qp(const eepacc_params* p, const eep_qp_io* io, cudaStream_t st)
{
    size_t sm = eep_qp_smem_bytes(io->n, io->m, io->sblk != nullptr);
    static size_t optin[EEP_MAX_DEV];                    // per device (eep_optin_smem)
    if (int e = eep_optin_smem(k3_qp_kernel, sm, optin)) return e;
    long long g = io->B;               // one CTA per vehicle: the hardware scheduler balances uneven solves
    if (io->todo_list && g > (io->todo_grid > 0 ? io->todo_grid : 148)) g = (io->todo_grid > 0 ? io->todo_grid : 148);      // fallback mode: a few persistent CTAs walk the (short) hand-over list
    if (g < 1) g = 1;
    if (g > 2147483647LL) g = 2147483647LL;
    k3_qp_kernel<<<(int)g, NT, sm, st>>>(*p, *io);
    return (int)cudaGetLastError();
}
