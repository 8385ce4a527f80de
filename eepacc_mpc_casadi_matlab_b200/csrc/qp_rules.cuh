// qp_rules.cuh -- decision rules shared by the two QP kernels (k3_qp.cu full space, k3r_qp.cu reduced space) and
// stated once in oracle/qp_oracle.c (rt_select, TIE_TOL, FLAT_TOL): the ratio tests are TWO-PASS (Harris):
//   pass 1   tmax = min_i (num_i + del_i) / den_i          del_i = round-off level of the numerator
//   pass 2   among the candidates with num_i / den_i <= tmax the LOWEST KEY wins (primal rows 2c+side before
//            dual rows 2mt+c: keyed by constraint index, never by working-set position, so the full-space and the
//            reduced-space factorisations -- whose positions differ -- take the same constraint)
//   a winner whose numerator is within del of zero takes a step of exactly zero.
// With these rules the sequence of working-set changes does not depend on summation order or FMA contraction at
// degenerate vertices (FBMPC standstill: identical rows at every stage tie exactly in theory and to 1e-16 in
// practice).  The eps-regularised LPs (BL/TV) keep exact comparisons (tie = 0), see qp_oracle.c.
// Included after the kernel's own block_max / lb_new / ub_new and its context type.
#pragma once

#define TIE_TOL   1.0e-11
#define FLAT_TOL  1.0e-9
#define RT_NONE   0x7fffffff

// two maxima in one pass (one pair of barriers)
template <class CT> __device__ __forceinline__ void block_max2(CT& C, double& a, double& b)
{
#ifdef EEP_HOST_EMU
    (void)C; (void)a; (void)b;
#elif (NT / 32) * 2 > 8
    a = block_max(C, a); b = block_max(C, b);      // the kernels' reduction scratch C.red holds 8 doubles
#else
    for (int o = 16; o > 0; o >>= 1) { a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o)); b = fmax(b, __shfl_xor_sync(0xffffffffu, b, o)); }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { C.red[threadIdx.x >> 5] = a; C.red[NT / 32 + (threadIdx.x >> 5)] = b; }
    __syncthreads();
    a = C.red[0]; b = C.red[NT / 32];
    for (int w = 1; w < NT / 32; ++w) { a = fmax(a, C.red[w]); b = fmax(b, C.red[NT / 32 + w]); }
#endif
}
template <class CT> __device__ __forceinline__ double block_min(CT& C, double v) { return -block_max(C, -v); }
// integer minimum (values < 2^52: exact as doubles)
template <class CT> __device__ __forceinline__ int block_imin(CT& C, int v) { return (int)(-block_max(C, -(double)v)); }

// primal candidate of the homotopy ratio test: row c, side s (0 lower / 1 upper) blocks before tau = 1?
template <class CT> __device__ __forceinline__ bool rt_primal(const CT& C, int c, int s, double dxmax, double& num, double& den)
{
    const double adx = C.Adx[c];
    const double rz = EPS_ROW * C.anorm[c] * dxmax;
    const double cur = s ? C.ubc[c] : C.lbc[c];
    if (isinf(cur)) return false;
    const double nw = s ? ub_new(C, c) : lb_new(C, c);
    const double db = isinf(nw) ? 0.0 : nw - cur;
    const double rate = s ? db - adx : adx - db;
    double sl = s ? cur - C.Ax[c] : C.Ax[c] - cur;
    if (sl < 0) sl = 0;
    if (rate < -(EPS_RATE * (1.0 + fabs(adx) + fabs(db)) + rz) && -rate - sl > FEAS_TOL * (1.0 + fabs(nw)) && sl < -rate) {
        num = sl; den = -rate;
        return true;
    }
    return false;
}
// dual candidate: multiplier yc of a row held at side st (-1 / +1) moving at rate dy.  del = tie * ynoise.
__device__ __forceinline__ bool rt_dual(int st, double yc, double dy, double thr, double del, double& num, double& den)
{
    const double dtol = fmax(FEAS_TOL * (1.0 + fabs(yc) + fabs(dy)), del);   // ends within dtol of zero at the full step: not blocking
    if (st < 0 && dy < -thr) { num = yc > 0 ? yc : 0; den = -dy; }
    else if (st > 0 && dy > thr) { num = yc < 0 ? -yc : 0; den = dy; }
    else return false;
    return num < den && den - num > dtol;
}
// primal candidate of the exchange ratio test along d (A d in C.Adx)
template <class CT> __device__ __forceinline__ bool rt_exch(const CT& C, int c, double dn, int& s, double& num, double& den)
{
    const double adx = C.Adx[c];
    const double rz = EPS_ROW * C.anorm[c] * dn;
    if (adx < -rz && !isinf(C.lbc[c])) { double sl = C.Ax[c] - C.lbc[c]; if (sl < 0) sl = 0; s = 0; num = sl; den = -adx; return true; }
    if (adx > rz && !isinf(C.ubc[c])) { double sl = C.ubc[c] - C.Ax[c]; if (sl < 0) sl = 0; s = 1; num = sl; den = adx; return true; }
    return false;
}

// ---- sequential Givens chains walked by ONE WARP (both kernels; MM(r, c) = packed triangle of the including file) ----
// The two wavefronts of a working-set change are chains: rotation k is only defined once rotation k-1 has passed over
// one column (row) of the triangle.  Walking them with the whole CTA costs two block barriers per rotation; here warp 0
// walks the chain alone, lane l owning the columns (rows) l, l+32, l+64, ..., and hands the rotation on by shuffle.
#ifndef EEP_HOST_EMU
// re-triangularise after the column rotations of an addition: row rotations k = 0..nR-2 on rows (k, k+1); the
// sub-diagonal entry (k+1, k) is in C.sub[k+1].  Call with every thread of warp 0.
template <class CT> __device__ __forceinline__ void warp_chain_rows(CT& C, int nR)
{
    const int lane = threadIdx.x & 31;
    for (int k = 0; k < nR - 1; ++k) {
        double c2 = 1.0, s2 = 0.0, rr = 0.0;
        if (lane == (k & 31)) {
            const double pa = MM(k, k), pb = C.sub[k + 1];
            rr = hypot(pa, pb);
            if (rr != 0.0) { c2 = pa / rr; s2 = pb / rr; MM(k, k) = rr; }
        }
        c2 = __shfl_sync(0xffffffffu, c2, k & 31); s2 = __shfl_sync(0xffffffffu, s2, k & 31); rr = __shfl_sync(0xffffffffu, rr, k & 31);
        if (rr != 0.0)
            for (int j = lane; j < nR; j += 32)
                if (j > k) {
                    const double u = MM(k, j), v = MM(k + 1, j);
                    MM(k, j) = c2 * u + s2 * v;
                    MM(k + 1, j) = -s2 * u + c2 * v;
                }
    }
}
// restore the triangle after a removal: column rotations r = rp .. nZ+1 on columns (r-1, r) over the rows nZ+1..r; the
// entry (r, r-1) is in C.sub[r].  Leaves the rotations in C.csA / C.snA for the Q update.  Call with every thread of warp 0.
template <class CT> __device__ __forceinline__ void warp_chain_cols(CT& C, int rp, int nZ)
{
    const int lane = threadIdx.x & 31;
    for (int r = rp; r > nZ; --r) {
        const int src = (rp - r) & 31;
        double cs = 1.0, sn = 0.0;
        if (lane == src) {
            const double a = C.sub[r], b = MM(r, r);
            const double rho = hypot(a, b);
            if (rho != 0.0) { cs = b / rho; sn = a / rho; }
            C.csA[r] = cs; C.snA[r] = sn;
            MM(r, r) = sn * a + cs * b;
        }
        cs = __shfl_sync(0xffffffffu, cs, src); sn = __shfl_sync(0xffffffffu, sn, src);
        for (int rr = rp - lane; rr > nZ; rr -= 32)
            if (rr < r) {
                const double u = MM(rr, r - 1), v = MM(rr, r);
                MM(rr, r - 1) = cs * u - sn * v;
                MM(rr, r) = sn * u + cs * v;
            }
    }
}
#endif
