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

#define NT 128
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
    // shared
    double *Q, *M, *sub;
    double *x, *gcur, *dx, *dg, *t1, *t2, *t3, *t4, *yW, *dyW, *dbw, *wv, *csA, *snA;
    double *Ax, *lbc, *ubc, *Adx, *anorm;
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

// ---- dense primitives; every routine ends with data visible after a __syncthreads() ----
__device__ void get_row(Ctx& C, int c, double* a)
{
    const int n = C.n, m = C.m;
    if (c < m) { for (int j = threadIdx.x; j < n; j += NT) a[j] = C.A[c + (size_t)j * m]; }
    else { for (int j = threadIdx.x; j < n; j += NT) a[j] = (j == c - m) ? 1.0 : 0.0; }
    __syncthreads();
}
__device__ void Hmul(Ctx& C, const double* v, double* out)
{
    const int n = C.n;
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
// solve M[lo:hi,lo:hi] x = b (back substitution); b is destroyed, x may not alias b
__device__ void usolve(Ctx& C, int lo, int hi, double* b, double* x)
{
    for (int r = hi - 1; r >= lo; --r) {
        __syncthreads();
        double xr = b[r] / MM(r, r);
        for (int i = lo + threadIdx.x; i < r; i += NT) b[i] -= MM(i, r) * xr;
        if (threadIdx.x == 0) x[r] = xr;
    }
    __syncthreads();
}
// solve M[lo:hi,lo:hi]' y = v (forward substitution); v destroyed
__device__ void tsolve(Ctx& C, int lo, int hi, double* v, double* y)
{
    for (int r = lo; r < hi; ++r) {
        __syncthreads();
        double yr = v[r] / MM(r, r);
        for (int i = r + 1 + threadIdx.x; i < hi; i += NT) v[i] -= MM(r, i) * yr;
        if (threadIdx.x == 0) y[r] = yr;
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
        // R phase B: re-triangularise (row rotations), wavefront over k
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
__device__ void kkt_step(Ctx& C, const double* dg, const double* dbw, double* dx, double* dyW)
{
    const int n = C.n, nZ = n - C.nA;
    double *p = C.t1, *u = C.t2, *t = C.t3;
    if (C.nA > 0) {
        for (int r = nZ + threadIdx.x; r < n; r += NT) t[r] = dbw[r];
        __syncthreads();
        usolve(C, nZ, n, t, p);
        Qmul(C, p, nZ, n, dx, false);
    } else {
        for (int i = threadIdx.x; i < n; i += NT) dx[i] = 0.0;
        __syncthreads();
    }
    Hmul(C, dx, u);
    for (int i = threadIdx.x; i < n; i += NT) u[i] += dg[i];
    __syncthreads();
    if (nZ > 0) {
        Qtmul(C, u, t, 0, nZ);
        for (int j = threadIdx.x; j < nZ; j += NT) t[j] = -t[j];
        __syncthreads();
        tsolve(C, 0, nZ, t, p);
        usolve(C, 0, nZ, p, t);
        Qmul(C, t, 0, nZ, p, false);            // p = Z q
        for (int i = threadIdx.x; i < n; i += NT) dx[i] += p[i];
        __syncthreads();
        Hmul(C, p, t);
        for (int i = threadIdx.x; i < n; i += NT) u[i] += t[i];
        __syncthreads();
    }
    if (C.nA > 0) {
        Qtmul(C, u, t, nZ, n);
        tsolve(C, nZ, n, t, dyW);
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
            if (c < m) s += C.A[c + (size_t)j * m] * yc;
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
            for (int i = threadIdx.x; i < n; i += NT) sgn -= (hz[i] + C.gcur[i]) * d[i];
            sgn = block_sum(C, sgn);
        }
        if (sgn < 0) { for (int i = threadIdx.x; i < n; i += NT) d[i] = -d[i]; __syncthreads(); }
        Amul(C, d, C.Adx);
        double dn = 0;
        for (int i = threadIdx.x; i < n; i += NT) dn += d[i] * d[i];
        dn = sqrt(block_sum(C, dn));
        double tmin = INFINITY; int key = 0x7fffffff;
        for (int cc = threadIdx.x; cc < mt; cc += NT) {
            if (C.st[cc] != 0 || C.anorm[cc] == 0.0) continue;
            double adx = C.Adx[cc];
            const double rz = EPS_ROW * C.anorm[cc] * dn;
            if (adx < -rz && !isinf(C.lbc[cc])) {
                double sl = C.Ax[cc] - C.lbc[cc]; if (sl < 0) sl = 0;
                double t = sl / (-adx);
                int k = 2 * cc;
                if (t < tmin || (t == tmin && k < key)) { tmin = t; key = k; }
            } else if (adx > rz && !isinf(C.ubc[cc])) {
                double sl = C.ubc[cc] - C.Ax[cc]; if (sl < 0) sl = 0;
                double t = sl / adx;
                int k = 2 * cc + 1;
                if (t < tmin || (t == tmin && k < key)) { tmin = t; key = k; }
            }
        }
        block_argmin(C, tmin, key);
        if (key == 0x7fffffff) return EEPACC_QP_UNBOUNDED;
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

extern __shared__ __align__(16) unsigned char k3_smem[];

__global__ void __launch_bounds__(NT) k3_qp_kernel(eep_qp_io io)
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
        C.Ax = p; p += mt; C.lbc = p; p += mt; C.ubc = p; p += mt; C.Adx = p; p += mt; C.anorm = p; p += mt;
        C.red = p; p += 8;
        C.ired = reinterpret_cast<int*>(p); p += 4;
        C.widx = reinterpret_cast<short*>(p); p += (n + 3) / 4 + 1;
        C.st = reinterpret_cast<signed char*>(p);
        C.tabu = C.st + mt;
    }
    for (long long veh = blockIdx.x; veh < io.B; veh += gridDim.x) {
        __syncthreads();
        C.H = io.H + veh * (size_t)n * n;
        C.g = io.g + veh * (size_t)n;
        C.A = io.A + veh * (size_t)m * n;
        C.lba = io.lba + veh * (size_t)m;
        C.uba = io.uba + veh * (size_t)m;
        C.lbx = io.lbx ? io.lbx + veh * (size_t)n : nullptr;
        C.ubx = io.ubx ? io.ubx + veh * (size_t)n : nullptr;
        int status = EEPACC_QP_OK, iter = 0;
        const int max_iter = 5 * (n + m);
        unsigned long long n_exch = 0, n_refac = 0;

        // Hessian scale and LP regularisation
        {
            double hd = 0, hm = 0;
            for (int e = threadIdx.x; e < n * n; e += NT) { double v = fabs(C.H[e]); hm = fmax(hm, v); if (e % (n + 1) == 0) hd = fmax(hd, v); }
            hd = block_max(C, hd);
            hm = block_max(C, hm);
            C.reg = (hm == 0.0) ? LP_REG : 0.0;
            C.hscale = fmax(hd, LP_REG);
            for (int r = threadIdx.x; r < m; r += NT) {
                double s = 0;
                for (int j = 0; j < n; ++j) s += fabs(C.A[r + (size_t)j * m]);
                C.anorm[r] = s;
            }
            for (int j = threadIdx.x; j < n; j += NT) C.anorm[m + j] = 1.0;
            for (int c = threadIdx.x; c < mt; c += NT) C.tabu[c] = 0;
            __syncthreads();
        }

        // ---------------- set-up ----------------
        const int was_warm = io.warm[veh];
        bool need_cold = !was_warm;
        for (int i = threadIdx.x; i < n; i += NT) C.x[i] = was_warm ? io.wx[veh * (size_t)n + i] : 0.0;
        if (!need_cold) {
            const int old_nA = io.wnA[veh];
            for (int c = threadIdx.x; c < mt; c += NT) C.st[c] = 0;
            for (int e = threadIdx.x; e < n * C.ldq; e += NT) C.Q[e] = 0.0;
            __syncthreads();
            for (int v = threadIdx.x; v < n; v += NT) MQ(v, v) = 1.0;
            C.nA = 0;
            __syncthreads();
            for (int k = 0; k < old_nA; ++k) {          // oldest first: stored column n-1-k
                const int rold = n - 1 - k;
                const int c = io.widx[veh * (size_t)n + rold];
                const int s = io.wside[veh * (size_t)n + rold];
                const double yv = io.wy[veh * (size_t)n + rold];
                double bnd = s < 0 ? lb_new(C, c) : ub_new(C, c);
                if (isinf(bnd)) continue;
                get_row(C, c, C.t2);
                const int nZ = n - C.nA;
                if (givens_add(C, c, s, C.t2, nZ, 0, true) == 0) {
                    if (threadIdx.x == 0) C.yW[nZ - 1] = yv;
                }
                __syncthreads();
            }
            n_refac++;
            if (refactor_R(C)) need_cold = true;
        }
        if (need_cold) cold_ws(C);
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
            kkt_step(C, C.dg, C.dbw, C.dx, C.dyW);
            Amul(C, C.dx, C.Adx);

            double dxmax = 0;
            for (int i = threadIdx.x; i < n; i += NT) dxmax = fmax(dxmax, fabs(C.dx[i]));
            dxmax = block_max(C, dxmax);
            double tau = 1.0; int key = 0x7fffffff;
            for (int c = threadIdx.x; c < mt; c += NT) {
                if (C.st[c] != 0 || C.anorm[c] == 0.0) continue;      // all-zero rows can never be added
                double adx = C.Adx[c];
                const double rz = EPS_ROW * C.anorm[c] * dxmax;
                if (!isinf(C.lbc[c])) {
                    double ln = lb_new(C, c);
                    double dlb = isinf(ln) ? 0.0 : ln - C.lbc[c];
                    double rate = adx - dlb;
                    double sl = C.Ax[c] - C.lbc[c]; if (sl < 0) sl = 0;
                    if (rate < -(EPS_RATE * (1.0 + fabs(adx) + fabs(dlb)) + rz) &&
                        -rate - sl > FEAS_TOL * (1.0 + fabs(ln))) {
                        double t = sl / (-rate);
                        int k = 2 * c;
                        if (t < tau || (t == tau && k < key && t < 1.0)) { tau = t; key = k; }
                    }
                }
                if (!isinf(C.ubc[c])) {
                    double un = ub_new(C, c);
                    double dub = isinf(un) ? 0.0 : un - C.ubc[c];
                    double rate = dub - adx;
                    double sl = C.ubc[c] - C.Ax[c]; if (sl < 0) sl = 0;
                    if (rate < -(EPS_RATE * (1.0 + fabs(adx) + fabs(dub)) + rz) &&
                        -rate - sl > FEAS_TOL * (1.0 + fabs(un))) {
                        double t = sl / (-rate);
                        int k = 2 * c + 1;
                        if (t < tau || (t == tau && k < key && t < 1.0)) { tau = t; key = k; }
                    }
                }
            }
            double dys = 0;
            for (int r = nZ + threadIdx.x; r < n; r += NT) dys = fmax(dys, fabs(C.dyW[r]));
            dys = block_max(C, dys);
            const double thr = EPS_RATE * (1.0 + dys);
            for (int r = nZ + threadIdx.x; r < n; r += NT) {
                int c = C.widx[r];
                if (lb_new(C, c) == ub_new(C, c)) continue;
                if (C.tabu[c]) continue;                 // no reversal inside a run of zero-length steps
                double yc = C.yW[r], dy = C.dyW[r];
                int k = 2 * mt + (n - 1 - r);
                if (C.st[c] < 0 && dy < -thr) {
                    double t = (yc > 0 ? yc : 0) / (-dy);
                    if (t < tau || (t == tau && k < key && t < 1.0)) { tau = t; key = k; }
                } else if (C.st[c] > 0 && dy > thr) {
                    double t = (yc < 0 ? -yc : 0) / dy;
                    if (t < tau || (t == tau && k < key && t < 1.0)) { tau = t; key = k; }
                }
            }
            block_argmin(C, tau, key);
            const bool full = !(tau < 1.0) || key == 0x7fffffff;
            if (full) tau = 1.0;
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
            if (full) break;
            if (++iter > max_iter) { status = EEPACC_QP_MAXITER; break; }
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
                    double theta = INFINITY; int kk = 0x7fffffff;
                    for (int r = nZc + threadIdx.x; r < n; r += NT) {
                        int c = C.widx[r];
                        if (lb_new(C, c) == ub_new(C, c)) continue;
                        double l = (side < 0) ? C.t2[r] : -C.t2[r];
                        if (fabs(l) <= 1e-11 * (1.0 + lmax)) continue;
                        double yc = C.yW[r];
                        int pos = n - 1 - r;
                        if (C.st[c] < 0 && l > 0) {
                            double th = (yc > 0 ? yc : 0) / l;
                            if (th < theta || (th == theta && pos < kk)) { theta = th; kk = pos; }
                        } else if (C.st[c] > 0 && l < 0) {
                            double th = (yc < 0 ? yc : 0) / l;
                            if (th < theta || (th == theta && pos < kk)) { theta = th; kk = pos; }
                        }
                    }
                    block_argmin(C, theta, kk);
                    if (kk == 0x7fffffff) { status = EEPACC_QP_INFEASIBLE; break; }
                    for (int r = nZc + threadIdx.x; r < n; r += NT) {
                        double l = (side < 0) ? C.t2[r] : -C.t2[r];
                        C.yW[r] -= theta * l;
                    }
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
                const int pos = key - 2 * mt;
                const int rp = n - 1 - pos;
                const int c = C.widx[rp];
                const int side = C.st[c];
                __syncthreads();
                remove_constraint(C, rp);
                int rc = resolve_pending(C, c, side, n_exch, zero_step);
                if (rc) { status = rc; break; }
            }
        }

        // ---------------- iterative refinement on the final working set ----------------
        if (status == EEPACC_QP_OK) {
            for (int pass = 0; pass < 2; ++pass) {
                const int nZ = n - C.nA;
                Hmul(C, C.x, C.t4);
                // rg = Hx + g - A_W' y ; rb = bound - a'x
                for (int j = threadIdx.x; j < n; j += NT) {
                    double s = C.t4[j] + C.g[j];
                    for (int r = n - 1; r >= nZ; --r) {
                        double yc = C.yW[r];
                        int c = C.widx[r];
                        if (c < m) s -= C.A[c + (size_t)j * m] * yc;
                        else if (c - m == j) s -= yc;
                    }
                    C.dg[j] = s;
                }
                for (int r = nZ + threadIdx.x; r < n; r += NT) {
                    int c = C.widx[r];
                    double ax = 0;
                    if (c < m) { for (int j = 0; j < n; ++j) ax += C.A[c + (size_t)j * m] * C.x[j]; }
                    else ax = C.x[c - m];
                    C.dbw[r] = (C.st[c] < 0 ? C.lbc[c] : C.ubc[c]) - ax;
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

extern "C" size_t eep_qp_smem_bytes(int n, int m)
{
    const int mt = n + m, ldq = n | 1;
    size_t d = (size_t)n * ldq + (size_t)n * (n + 1) / 2 + (n + 1) + 12 * (size_t)n + 2 * (size_t)(n + 1)
             + 5 * (size_t)mt + 8 + 4 + ((n + 3) / 4 + 1);
    return d * 8 + 2 * (size_t)mt + 16;
}

extern "C" int eep_launch_qp(const eep_qp_io* io, cudaStream_t st)
{
    size_t sm = eep_qp_smem_bytes(io->n, io->m);
    static size_t attr_for = 0;
    if (sm > attr_for) {
        cudaError_t e = cudaFuncSetAttribute(k3_qp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return (int)e;
        attr_for = sm;
    }
    long long g = io->B;
    const long long cap = 148LL * 4;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    k3_qp_kernel<<<(int)g, NT, sm, st>>>(*io);
    return (int)cudaGetLastError();
}
