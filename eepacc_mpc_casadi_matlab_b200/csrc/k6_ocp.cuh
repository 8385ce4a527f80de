// k6_ocp.cuh -- K6: stage-structured (OCP) interior-point QP solver, ONE VEHICLE PER THREAD.
//
// Stands where the reference selects its sparse / OCP-structured solvers instead of condensing + dense qpOASES:
//   OPTsettings.solverToUse = 0 (sparse qpOASES) / 2 (HPIPM, a Riccati-based interior-point method for OCP QPs),
//   MC/Settings.m:112, RunOpt_FBMPC.m:96-121, RunOpt_ABMPC.m:87-117, CreateQP_FB.m:219-308 (dynamics kept as stage
//   equalities, previous-stage quantities carried as extra states), CreateQP_AB.m:174-235.
// SURVEY.md 8(f)-3.  HPIPM itself is a binary-only CasADi plugin in the reference tree; this is the published
// algorithm class (Mehrotra predictor-corrector on the stage-structured KKT system, backward Riccati recursion;
// Frison & Diehl, "HPIPM", IFAC 2020; Rao, Wright & Rawlings, JOTA 1998), not a port.
//
// The QP solved is EXACTLY the one the dense path solves (K2's structured block, the packed row table and the
// shifted bounds lba/uba are shared with K3r), written in the deviation coordinates of the free response:
//   states  x_k = (ds_k, dv_k, e_k),  e_k = [dv_(k-1), driven controls of stage k-1]   (carried, as HPIPM's
//                                                                       vprev/Fmprev/Fbprev of CreateQP_FB.m:226-246)
//   ds_(k+1) = ds_k + T dv_k + M0 sum(driven_k),   dv_(k+1) = alpha_k dv_k + M1 sum(driven_k),   x_0 = 0
//   cost    sum_k 1/2 c_k' D_k c_k + c_k' C_k e_k + 1/2 sum Dg u^2 + tv_k dv_k + qu_k' u_k,  c_k = [dv_k, driven_k]
//   rows    lba <= a_r' (x_k, u_k) <= uba   (every row touches one stage: eepacc_rows.cuh)
// so first-stage controls, objective and closed-loop states can be compared with the dense path directly
// (tests/test_ocp_host_emu.py, tests/test_gpu_ocp.py).  Where the optimum is not unique (FBMPC at standstill: the
// Fm/Fb split is a flat direction of the objective) an interior-point method returns the analytic centre of the
// optimal face, the active-set path a vertex: the objective and the states agree, the split does not.
//
// Why one vehicle per thread: a Newton step costs ~4e4 flop, strictly sequential along the 20 stages, on 11 x 11
// blocks -- too small for a CTA (K3r spends its time in barriers at n = 40).  Here nothing is shared and nothing
// synchronises; all vehicles of a launch are in flight at once, per-vehicle work arrays are laid out [item][B] so
// that the 32 vehicles of a warp read consecutive addresses (HBM/L2-bound streaming of t, lambda; the stage blocks
// live in registers / local memory).
#pragma once
#include "eepacc_dev.cuh"
#include "eepacc_kernels.h"
#ifdef K6_DEBUG_CHECK
#include <cstdio>
#endif

#ifndef K6_MAXIT
#define K6_MAXIT 80
#endif
#define K6_BIGB 1.0e19

namespace k6 {
#ifdef K6_DEBUG_CHECK
static double k6dbg_F[EEP_N + 1][16][16], k6dbg_f[EEP_N + 1][16];
#endif

template <int NU, int NB>
struct Dim {
    static constexpr int NE = 1 + NB, NX = 2 + NE, NZ = NX + NU, NL = NU * (NU + 1) / 2;
};

// doubles of per-vehicle work space (each "item" is one double per vehicle, stride = batch)
__host__ __device__ inline long long work_items(int n, int m, int nu, int nb)
{
    const int NH = EEP_N, nx = 3 + nb;
    return 6LL * m + 3LL * n + 3LL * 2 * (NH + 1) + (long long)NH * (nu * (nu + 1) / 2 + nu * nx + nu)
           + (NH + 1 + n) + (EEP_SB_DG + 4 * NH - EEP_SB_VEST) + NH;
}

struct SparseRow { int nnz; int idx[8]; double val[8]; };
// K2's structured block seen through the [item][B] work space (index = offset inside the AoS block)
struct SbView {
    const double* base; long long ws;
    __device__ __forceinline__ double operator[](int i) const { return base[(long long)i * ws]; }
};

// stage-local coefficient vector of packed row c (stage k): indices 0..NX-1 states, NX.. controls
template <int NU, int NB>
__device__ __forceinline__ void load_row(unsigned long long w, const double* __restrict__ rowc, const SbView& sb,
                                         int k, SparseRow& r)
{
    constexpr int NX = Dim<NU, NB>::NX;
    const int kst = (int)(signed char)(w & 0xff), aux = (int)((w >> 8) & 0xff);
    r.nnz = 0;
    if (kst >= 0) {
        if (aux & 128) { r.idx[r.nnz] = 0; r.val[r.nnz] = 1.0; ++r.nnz; }
        const int id = aux & 127;
        const double b = rowc[64 + id];
        const double cV = rowc[32 + id] + (b != 0.0 ? b * sb[EEP_SB_VEST + kst] : 0.0);
        if (cV != 0.0) { r.idx[r.nnz] = 1; r.val[r.nnz] = cV; ++r.nnz; }
    } else if (kst == -2) {
        for (int c = 0; c < NB; ++c) { r.idx[r.nnz] = 3 + c; r.val[r.nnz] = -1.0; ++r.nnz; }
    }
    for (int e = 0; e < 3; ++e) {
        const int j = (int)((w >> (16 + 8 * e)) & 0xff);
        if (j == 255) continue;
        const int id = (int)((w >> (40 + 8 * e)) & 0xff);
        r.idx[r.nnz] = NX + (j - k * NU); r.val[r.nnz] = rowc[id]; ++r.nnz;
    }
}

struct Args {
    const double* sb;           // this vehicle's structured block [EEP_SBLK]
    const double* alpha;        // [N] A_k(2,2)
    const double* qz;           // [N + 1 + n]: tv_k (k = 0..N), qu
    const double* lba;          // [m]
    const double* uba;          // [m]
    const unsigned long long* rowtab;
    const double* rowc;
    double* work;               // first item of this vehicle (work + veh)
    long long ws;               // stride between items (batch size)
    int m;
    double T, M0, M1;
    double mu0, eps_mu, eps_p, eps_d;
    double tau0;                // initial slacks are at least tau0 x the row's coefficient scale
    double gamma;               // width of the central-path neighbourhood the corrector step must stay in
    double eps_accept;          // a solve whose factorisation breaks down below this complementarity is accepted as it stands
    int maxit;
    int warm;                   // 1: u, t, lambda in the work space hold the previous step's solution
};

#define K6W(off) A.work[(long long)(off) * A.ws]

// one thread = one vehicle.  Returns status (0 solved, 1 iteration limit, 2 numerical breakdown); out: x[n], cost, iterations
template <int NU, int NB>
__device__ int solve(const Args& A, double* __restrict__ xout, double& cost_out, int& iter_out)
{
    using D = Dim<NU, NB>;
    constexpr int NH = EEP_N, NE = D::NE, NX = D::NX, NZ = D::NZ, NL = D::NL;
    constexpr int n = NH * NU;
    const int m = A.m, rps = m / NH, nterm = m - rps * NH;
    // work-space offsets (items)
    const int oTL = 0, oLL = m, oTU = 2 * m, oLU = 3 * m, oU = 4 * m, oDUA = oU + n, oDU = oDUA + n;
    const int oSX = oDU + n, oDSXA = oSX + 2 * (NH + 1), oDSX = oDSXA + 2 * (NH + 1);
    const int oL = oDSX + 2 * (NH + 1), oK = oL + NH * NL, oKF = oK + NH * NU * NX;
    // the vehicle's operands, copied once from the AoS blocks K2 writes into the [item][B] layout (coalesced from then on)
    const int oLB = oKF + NH * NU, oUB = oLB + m, oQZ = oUB + m, oSB = oQZ + (NH + 1 + n) - EEP_SB_VEST;
    const int oAL = oSB + EEP_SB_DG + 4 * NH;
    for (int c = 0; c < m; ++c) { K6W(oLB + c) = A.lba[c]; K6W(oUB + c) = A.uba[c]; }
    for (int j = 0; j < NH + 1 + n; ++j) K6W(oQZ + j) = A.qz[j];
    for (int j = EEP_SB_VEST; j < EEP_SB_DG + 4 * NH; ++j) K6W(oSB + j) = A.sb[j];
    for (int k = 0; k < NH; ++k) K6W(oAL + k) = A.alpha[k];
    const SbView sb{ A.work + (long long)oSB * A.ws, A.ws };

    // ---- initial point ----
    int nside = 0;
    if (!A.warm) {
        for (int j = 0; j < n; ++j) K6W(oU + j) = 0.0;
        for (int j = 0; j < 2 * (NH + 1); ++j) K6W(oSX + j) = 0.0;
    }
    for (int k = 0; k <= NH; ++k) {
        const int c0 = (k < NH) ? k * rps : NH * rps, c1 = (k < NH) ? c0 + rps : c0 + nterm;
        for (int c = c0; c < c1; ++c) {
            SparseRow r; load_row<NU, NB>(A.rowtab[c], A.rowc, sb, k, r);
            double nrm = 0.0;
            for (int e = 0; e < r.nnz; ++e) nrm = fmax(nrm, fabs(r.val[e]));
            // a row whose only entries sit on the fixed initial state is a constant: nothing to solve for
            bool live = false;
            for (int e = 0; e < r.nnz; ++e) if (k > 0 || r.idx[e] >= NX) live = true;
            const double lb = K6W(oLB + c), ub = K6W(oUB + c);
            const bool hl = live && lb > -K6_BIGB, hu = live && ub < K6_BIGB;
            double az = 0.0;
            if (A.warm && live) {
                for (int e = 0; e < r.nnz; ++e) {
                    const int i = r.idx[e];
                    double z;
                    if (i == 0) z = K6W(oSX + 2 * k);
                    else if (i == 1) z = K6W(oSX + 2 * k + 1);
                    else if (i == 2) z = (k > 0) ? K6W(oSX + 2 * (k - 1) + 1) : 0.0;
                    else if (i < NX) z = (k > 0) ? K6W(oU + (k - 1) * NU + (i - 3)) : 0.0;
                    else z = K6W(oU + k * NU + (i - NX));
                    az += r.val[e] * z;
                }
            }
            if (hl) {
                const double t = fmax(az - lb, A.tau0 * nrm);
                K6W(oTL + c) = t; K6W(oLL + c) = A.mu0 / t; ++nside;
            } else { K6W(oTL + c) = 1.0; K6W(oLL + c) = 0.0; }
            if (hu) {
                const double t = fmax(ub - az, A.tau0 * nrm);
                K6W(oTU + c) = t; K6W(oLU + c) = A.mu0 / t; ++nside;
            } else { K6W(oTU + c) = 1.0; K6W(oLU + c) = 0.0; }
        }
    }
    if (nside == 0) nside = 1;

    double nu_res = 1.0;            // factor by which the (linear) stationarity residual has shrunk
    int status = 1, it = 0;
    double mu = 0.0;
    for (; it < A.maxit; ++it) {
        // ================= pass 0: mu and the primal residual =================
        double smu = 0.0, rpmax = 0.0;
        for (int k = 0; k <= NH; ++k) {
            double zk[NZ];
            zk[0] = K6W(oSX + 2 * k); zk[1] = K6W(oSX + 2 * k + 1);
            zk[2] = (k > 0) ? K6W(oSX + 2 * (k - 1) + 1) : 0.0;
            for (int c = 0; c < NB; ++c) zk[3 + c] = (k > 0) ? K6W(oU + (k - 1) * NU + c) : 0.0;
            for (int c = 0; c < NU; ++c) zk[NX + c] = (k < NH) ? K6W(oU + k * NU + c) : 0.0;
            const int c0 = (k < NH) ? k * rps : NH * rps, c1 = (k < NH) ? c0 + rps : c0 + nterm;
            for (int c = c0; c < c1; ++c) {
                const double ll = K6W(oLL + c), lu = K6W(oLU + c);
                if (ll == 0.0 && lu == 0.0) continue;
                SparseRow r; load_row<NU, NB>(A.rowtab[c], A.rowc, sb, k, r);
                double az = 0.0, nrm = 0.0;
                for (int e = 0; e < r.nnz; ++e) { az += r.val[e] * zk[r.idx[e]]; nrm = fmax(nrm, fabs(r.val[e])); }
                if (ll != 0.0) { const double t = K6W(oTL + c); smu += t * ll; rpmax = fmax(rpmax, fabs(az - t - K6W(oLB + c)) / nrm); }
                if (lu != 0.0) { const double t = K6W(oTU + c); smu += t * lu; rpmax = fmax(rpmax, fabs(-az - t + K6W(oUB + c)) / nrm); }
            }
        }
        mu = smu / nside;
#ifdef K6_DEBUG_CHECK
        printf("it %d: mu %.3e rpmax %.3e nu %.3e sigma(prev) %.3e\n", it, mu, rpmax, nu_res, 0.0);
#endif
        if (mu <= A.eps_mu && rpmax <= A.eps_p && nu_res <= A.eps_d) { status = 0; break; }
        if (!(mu == mu) || !(rpmax == rpmax)) { status = 2; break; }

        double sigma = 0.0, apA = 1.0, adA = 1.0;
        double alpha_step = 1.0;
        bool breakdown = false;     // a Riccati pivot was lost (barrier weights beyond what FP64 resolves, or negative curvature)
        for (int pass = 0; pass < 2; ++pass) {
            // ================= backward sweep: (pass 0) factorise + affine rhs, (pass 1) corrector rhs =================
            double P[NX][NX], pv[NX];
            for (int a = 0; a < NX; ++a) { pv[a] = 0.0; for (int b = 0; b < NX; ++b) P[a][b] = 0.0; }
            for (int k = NH; k >= 0; --k) {
                double F[NZ][NZ], f[NZ], zk[NZ], dk[NZ];
                for (int a = 0; a < NZ; ++a) { f[a] = 0.0; dk[a] = 0.0; if (pass == 0) for (int b = 0; b < NZ; ++b) F[a][b] = 0.0; }
                zk[0] = K6W(oSX + 2 * k); zk[1] = K6W(oSX + 2 * k + 1);
                zk[2] = (k > 0) ? K6W(oSX + 2 * (k - 1) + 1) : 0.0;
                for (int c = 0; c < NB; ++c) zk[3 + c] = (k > 0) ? K6W(oU + (k - 1) * NU + c) : 0.0;
                for (int c = 0; c < NU; ++c) zk[NX + c] = (k < NH) ? K6W(oU + k * NU + c) : 0.0;
                if (pass == 1) {
                    dk[0] = K6W(oDSXA + 2 * k); dk[1] = K6W(oDSXA + 2 * k + 1);
                    dk[2] = (k > 0) ? K6W(oDSXA + 2 * (k - 1) + 1) : 0.0;
                    for (int c = 0; c < NB; ++c) dk[3 + c] = (k > 0) ? K6W(oDUA + (k - 1) * NU + c) : 0.0;
                    for (int c = 0; c < NU; ++c) dk[NX + c] = (k < NH) ? K6W(oDUA + k * NU + c) : 0.0;
                }
                // ---- stage cost: Hessian blocks and gradient at z ----
                if (k < NH) {
                    int ci[1 + NB], ei[1 + NB];
                    ci[0] = 1; ei[0] = 2;
                    for (int c = 0; c < NB; ++c) { ci[1 + c] = NX + c; ei[1 + c] = 3 + c; }
                    for (int a = 0; a <= NB; ++a) {
                        double ga = 0.0;
                        for (int b = 0; b <= NB; ++b) {
                            const double d = sb[EEP_SB_DC + k * 9 + a * 3 + b], cc = sb[EEP_SB_CC + k * 9 + a * 3 + b];
                            if (pass == 0) { F[ci[a]][ci[b]] += d; F[ci[a]][ei[b]] += cc; F[ei[b]][ci[a]] += cc; }
                            ga += d * zk[ci[b]] + cc * zk[ei[b]];
                            f[ei[b]] += cc * zk[ci[a]];
                        }
                        f[ci[a]] += ga;
                    }
                    for (int c = NB; c < NU; ++c) {
                        const double d = sb[EEP_SB_DG + k * 4 + (c - NB)];
                        if (pass == 0) F[NX + c][NX + c] += d;
                        f[NX + c] += d * zk[NX + c];
                    }
                    for (int c = 0; c < NU; ++c) f[NX + c] += K6W(oQZ + NH + 1 + k * NU + c);
                }
                f[1] += K6W(oQZ + k);
                // ---- rows of the stage: barrier terms ----
                const int c0 = (k < NH) ? k * rps : NH * rps, c1 = (k < NH) ? c0 + rps : c0 + nterm;
                for (int c = c0; c < c1; ++c) {
                    const double ll = K6W(oLL + c), lu = K6W(oLU + c);
                    if (ll == 0.0 && lu == 0.0) continue;
                    SparseRow r; load_row<NU, NB>(A.rowtab[c], A.rowc, sb, k, r);
                    double az = 0.0, adz = 0.0;
                    for (int e = 0; e < r.nnz; ++e) { az += r.val[e] * zk[r.idx[e]]; adz += r.val[e] * dk[r.idx[e]]; }
                    double w = 0.0, cf = 0.0;
                    if (ll != 0.0) {
                        const double t = K6W(oTL + c), rp = az - t - K6W(oLB + c);
                        double corr = 0.0;
                        if (pass == 1) {
                            const double dtA = adz + rp, dlA = (-t * ll - ll * dtA) / t;
                            corr = apA * adA * dtA * dlA;
                        }
                        w += ll / t;
                        cf -= (sigma * mu - corr - ll * rp) / t;
                    }
                    if (lu != 0.0) {
                        const double t = K6W(oTU + c), rp = -az - t + K6W(oUB + c);
                        double corr = 0.0;
                        if (pass == 1) {
                            const double dtA = -adz + rp, dlA = (-t * lu - lu * dtA) / t;
                            corr = apA * adA * dtA * dlA;
                        }
                        w += lu / t;
                        cf += (sigma * mu - corr - lu * rp) / t;
                    }
                    for (int e = 0; e < r.nnz; ++e) {
                        f[r.idx[e]] += cf * r.val[e];
                        if (pass == 0) {
                            const double wv = w * r.val[e];
                            for (int e2 = 0; e2 < r.nnz; ++e2) F[r.idx[e]][r.idx[e2]] += wv * r.val[e2];
                        }
                    }
                }
#ifdef K6_DEBUG_CHECK
                for (int a = 0; a < NZ; ++a) { k6dbg_f[k][a] = f[a]; if (pass == 0) for (int b = 0; b < NZ; ++b) k6dbg_F[k][a][b] = F[a][b]; }
#endif
                if (k == NH) {
                    // terminal stage: only (ds_N, dv_N) carry cost / rows; e_N is not used
                    for (int a = 0; a < NX; ++a) { pv[a] = f[a]; if (pass == 0) for (int b = 0; b < NX; ++b) P[a][b] = F[a][b]; }
                    continue;
                }
                // ---- cost-to-go through the dynamics: [A B]' P [A B], [A B]' p ----
                // x+ rows: ds+ = ds + T dv + M0 sum(drv); dv+ = alpha dv + M1 sum(drv); e0+ = dv; e(1+c)+ = u_c
                const double al = K6W(oAL + k);
                {
                    // g = [A B]' pv
                    double drvp = A.M0 * pv[0] + A.M1 * pv[1];
                    f[0] += pv[0];
                    f[1] += A.T * pv[0] + al * pv[1] + pv[2];
                    for (int c = 0; c < NB; ++c) f[NX + c] += drvp + pv[3 + c];
                }
                if (pass == 0) {
                    // PA = P [A B]  (NX x NZ), then F += [A B]' PA
                    double PA[NX][NZ];
                    for (int a = 0; a < NX; ++a) {
                        for (int b = 0; b < NZ; ++b) PA[a][b] = 0.0;
                        PA[a][0] = P[a][0];
                        PA[a][1] = A.T * P[a][0] + al * P[a][1] + P[a][2];
                        const double drv = A.M0 * P[a][0] + A.M1 * P[a][1];
                        for (int c = 0; c < NB; ++c) PA[a][NX + c] = drv + P[a][3 + c];
                    }
                    for (int b = 0; b < NZ; ++b) {
                        F[0][b] += PA[0][b];
                        F[1][b] += A.T * PA[0][b] + al * PA[1][b] + PA[2][b];
                        const double drv = A.M0 * PA[0][b] + A.M1 * PA[1][b];
                        for (int c = 0; c < NB; ++c) F[NX + c][b] += drv + PA[3 + c][b];
                    }
                    // ---- Cholesky of Ruu (lower, packed), tiny relative regularisation on the flat directions ----
                    double Lc[NL];
                    bool bad = false;
                    for (int a = 0; a < NU; ++a) {
                        for (int b = 0; b <= a; ++b) {
                            double s = F[NX + a][NX + b];
                            for (int q = 0; q < b; ++q) s -= Lc[a * (a + 1) / 2 + q] * Lc[b * (b + 1) / 2 + q];
                            if (a == b) {
                                // a pivot lost to round-off (flat direction, e.g. the FBMPC force split far from its bounds)
                                // is replaced by a pivot-relative floor: the step in that direction becomes a damped one
                                const double floor_ = 1e-13 * fabs(F[NX + a][NX + a]);
                                if (!(s > floor_)) { bad = bad || !(s > 0.0); s = fabs(F[NX + a][NX + a]) + 1.0; }
                                Lc[a * (a + 1) / 2 + a] = sqrt(s);
                            } else Lc[a * (a + 1) / 2 + b] = s / Lc[b * (b + 1) / 2 + b];
                        }
                    }
                    if (bad) breakdown = true;
#ifdef K6_DEBUG_CHECK
                    if (bad) { printf("   stage %d: lost pivot; Ruu diag:", k); for (int a = 0; a < NU; ++a) printf(" %.3e", F[NX + a][NX + a]); printf("\n"); }
#endif
                    for (int q = 0; q < NL; ++q) K6W(oL + k * NL + q) = Lc[q];
                    // K = -Ruu^-1 Sux  (column by column), P_k = Qxx + Sux' K
                    for (int b = 0; b < NX; ++b) {
                        double y[NU];
                        for (int a = 0; a < NU; ++a) {
                            double s = -F[NX + a][b];
                            for (int q = 0; q < a; ++q) s -= Lc[a * (a + 1) / 2 + q] * y[q];
                            y[a] = s / Lc[a * (a + 1) / 2 + a];
                        }
                        for (int a = NU - 1; a >= 0; --a) {
                            double s = y[a];
                            for (int q = a + 1; q < NU; ++q) s -= Lc[q * (q + 1) / 2 + a] * y[q];
                            y[a] = s / Lc[a * (a + 1) / 2 + a];
                        }
                        for (int a = 0; a < NU; ++a) K6W(oK + (k * NU + a) * NX + b) = y[a];
                        for (int a2 = 0; a2 < NX; ++a2) {
                            double s = F[a2][b];
                            for (int a = 0; a < NU; ++a) s += F[NX + a][a2] * y[a];
                            P[a2][b] = s;
                        }
                    }
                    // symmetrise (round-off)
                    for (int a = 0; a < NX; ++a) for (int b = 0; b < a; ++b) { const double s = 0.5 * (P[a][b] + P[b][a]); P[a][b] = s; P[b][a] = s; }
                }
                // ---- feed-forward kff = -Ruu^-1 ru and p_k = qx + K' ru ----
                {
                    double y[NU];
                    for (int a = 0; a < NU; ++a) {
                        double s = -f[NX + a];
                        for (int q = 0; q < a; ++q) s -= K6W(oL + k * NL + a * (a + 1) / 2 + q) * y[q];
                        y[a] = s / K6W(oL + k * NL + a * (a + 1) / 2 + a);
                    }
                    for (int a = NU - 1; a >= 0; --a) {
                        double s = y[a];
                        for (int q = a + 1; q < NU; ++q) s -= K6W(oL + k * NL + q * (q + 1) / 2 + a) * y[q];
                        y[a] = s / K6W(oL + k * NL + a * (a + 1) / 2 + a);
                    }
                    for (int a = 0; a < NU; ++a) K6W(oKF + k * NU + a) = y[a];
                    for (int b = 0; b < NX; ++b) {
                        double s = f[b];
                        for (int a = 0; a < NU; ++a) s += K6W(oK + (k * NU + a) * NX + b) * f[NX + a];
                        pv[b] = s;
                    }
                }
            }
            if (pass == 0 && breakdown && mu <= A.eps_accept && rpmax <= A.eps_p) break;   // converged as far as FP64 goes
            // ================= forward sweep: the step =================
            const int oD = pass == 0 ? oDUA : oDU, oS = pass == 0 ? oDSXA : oDSX;
            {
                double xk[NX];
                for (int a = 0; a < NX; ++a) xk[a] = 0.0;
                K6W(oS) = 0.0; K6W(oS + 1) = 0.0;
                for (int k = 0; k < NH; ++k) {
                    double uk[NU];
                    for (int a = 0; a < NU; ++a) {
                        double s = K6W(oKF + k * NU + a);
                        for (int b = 0; b < NX; ++b) s += K6W(oK + (k * NU + a) * NX + b) * xk[b];
                        uk[a] = s;
                        K6W(oD + k * NU + a) = s;
                    }
                    double drv = 0.0;
                    for (int c = 0; c < NB; ++c) drv += uk[c];
                    const double sn = xk[0] + A.T * xk[1] + A.M0 * drv, vn = K6W(oAL + k) * xk[1] + A.M1 * drv;
                    xk[2] = xk[1];
                    for (int c = 0; c < NB; ++c) xk[3 + c] = uk[c];
                    xk[0] = sn; xk[1] = vn;
                    K6W(oS + 2 * (k + 1)) = sn; K6W(oS + 2 * (k + 1) + 1) = vn;
                }
            }
#ifdef K6_DEBUG_CHECK
            {   // directional derivative of the Newton subproblem along random feasible directions must vanish
                unsigned rs = 12345u + 77u * it + pass;
                double worst = 0.0, scale = 0.0;
                for (int trial = 0; trial < 4; ++trial) {
                    double wu[NH * NU], ws_[NH + 1], wv_[NH + 1];
                    for (int j = 0; j < NH * NU; ++j) { rs = rs * 1664525u + 1013904223u; wu[j] = ((rs >> 8) & 0xffff) / 65536.0 - 0.5; }
                    ws_[0] = wv_[0] = 0.0;
                    for (int k = 0; k < NH; ++k) { double drv = 0; for (int c = 0; c < NB; ++c) drv += wu[k * NU + c];
                        ws_[k + 1] = ws_[k] + A.T * wv_[k] + A.M0 * drv; wv_[k + 1] = A.alpha[k] * wv_[k] + A.M1 * drv; }
                    double acc = 0.0, sc = 0.0;
                    for (int k = 0; k <= NH; ++k) {
                        double wk[NZ], dk[NZ];
                        wk[0] = ws_[k]; wk[1] = wv_[k]; wk[2] = k > 0 ? wv_[k - 1] : 0.0;
                        dk[0] = K6W(oS + 2 * k); dk[1] = K6W(oS + 2 * k + 1); dk[2] = k > 0 ? K6W(oS + 2 * (k - 1) + 1) : 0.0;
                        for (int c = 0; c < NB; ++c) { wk[3 + c] = k > 0 ? wu[(k - 1) * NU + c] : 0.0; dk[3 + c] = k > 0 ? K6W(oD + (k - 1) * NU + c) : 0.0; }
                        for (int c = 0; c < NU; ++c) { wk[NX + c] = k < NH ? wu[k * NU + c] : 0.0; dk[NX + c] = k < NH ? K6W(oD + k * NU + c) : 0.0; }
                        for (int a = 0; a < NZ; ++a) {
                            double g = k6dbg_f[k][a];
                            for (int b = 0; b < NZ; ++b) g += k6dbg_F[k][a][b] * dk[b];
                            acc += wk[a] * g; sc += fabs(wk[a] * g);
                        }
                    }
                    if (fabs(acc) > worst) { worst = fabs(acc); scale = sc; }
                }
                printf("   it %d pass %d: newton residual along random directions %.3e (scale %.3e)\n", it, pass, worst, scale);
            }
#endif
            // ================= rows: step lengths (and mu_aff after the predictor) =================
            double ap = 1.0, ad = 1.0, a_try = 1.0;
            for (int rep = 0; rep < (pass == 0 ? 2 : 5); ++rep) {
                // rep 0: largest feasible step; rep 1 (predictor): complementarity after that step; rep >= 1 (corrector):
                // the step is halved while it would leave the neighbourhood t_i lambda_i >= gamma mu of the central path
                // (a pair driven to the boundary far ahead of the others jams every later step)
                double smuA = 0.0, minprod = 1e300;
                const double sp = (pass == 0) ? ap : a_try, sd = (pass == 0) ? ad : a_try;
                for (int k = 0; k <= NH; ++k) {
                    double zk[NZ], dk[NZ], dA[NZ];
                    zk[0] = K6W(oSX + 2 * k); zk[1] = K6W(oSX + 2 * k + 1);
                    zk[2] = (k > 0) ? K6W(oSX + 2 * (k - 1) + 1) : 0.0;
                    dk[0] = K6W(oS + 2 * k); dk[1] = K6W(oS + 2 * k + 1);
                    dk[2] = (k > 0) ? K6W(oS + 2 * (k - 1) + 1) : 0.0;
                    for (int c = 0; c < NB; ++c) { zk[3 + c] = (k > 0) ? K6W(oU + (k - 1) * NU + c) : 0.0; dk[3 + c] = (k > 0) ? K6W(oD + (k - 1) * NU + c) : 0.0; }
                    for (int c = 0; c < NU; ++c) { zk[NX + c] = (k < NH) ? K6W(oU + k * NU + c) : 0.0; dk[NX + c] = (k < NH) ? K6W(oD + k * NU + c) : 0.0; }
                    if (pass == 1) {
                        dA[0] = K6W(oDSXA + 2 * k); dA[1] = K6W(oDSXA + 2 * k + 1);
                        dA[2] = (k > 0) ? K6W(oDSXA + 2 * (k - 1) + 1) : 0.0;
                        for (int c = 0; c < NB; ++c) dA[3 + c] = (k > 0) ? K6W(oDUA + (k - 1) * NU + c) : 0.0;
                        for (int c = 0; c < NU; ++c) dA[NX + c] = (k < NH) ? K6W(oDUA + k * NU + c) : 0.0;
                    }
                    const int c0 = (k < NH) ? k * rps : NH * rps, c1 = (k < NH) ? c0 + rps : c0 + nterm;
                    for (int c = c0; c < c1; ++c) {
                        const double ll = K6W(oLL + c), lu = K6W(oLU + c);
                        if (ll == 0.0 && lu == 0.0) continue;
                        SparseRow r; load_row<NU, NB>(A.rowtab[c], A.rowc, sb, k, r);
                        double az = 0.0, adz = 0.0, adAz = 0.0;
                        for (int e = 0; e < r.nnz; ++e) { az += r.val[e] * zk[r.idx[e]]; adz += r.val[e] * dk[r.idx[e]]; }
                        if (pass == 1) for (int e = 0; e < r.nnz; ++e) adAz += r.val[e] * dA[r.idx[e]];
                        for (int side = 0; side < 2; ++side) {
                            const double lam = side ? lu : ll;
                            if (lam == 0.0) continue;
                            const double t = K6W((side ? oTU : oTL) + c);
                            const double sg = side ? -1.0 : 1.0;
                            const double rp = sg * az - t - (side ? -K6W(oUB + c) : K6W(oLB + c));
                            double rc = sigma * mu - t * lam;
                            if (pass == 1) {
                                const double dtA = sg * adAz + rp, dlA = (-t * lam - lam * dtA) / t;
                                rc -= apA * adA * dtA * dlA;
                            }
                            const double dt = sg * adz + rp, dl = (rc - lam * dt) / t;
                            if (rep == 0) {
                                if (dt < 0.0) ap = fmin(ap, -t / dt);
                                if (dl < 0.0) ad = fmin(ad, -lam / dl);
                            } else { const double pr = (t + sp * dt) * (lam + sd * dl); smuA += pr; minprod = fmin(minprod, pr); }
                        }
                    }
                }
                if (pass == 1) {
                    if (rep == 0) { const double a = fmin(ap, ad); a_try = (a < 1.0) ? 0.995 * a : 1.0; }
                    else if (minprod >= A.gamma * (smuA / nside) || rep == 4) break;
                    else a_try *= 0.5;
                } else if (rep == 1) {
                    const double muA = smuA / nside;
                    double q = muA / mu;
                    if (q < 0.0) q = 0.0;
                    sigma = fmin(1.0, q * q * q);
                    apA = ap; adA = ad;
                }
            }
            if (pass == 1) {
                alpha_step = a_try;
#ifdef K6_DEBUG_CHECK
                printf("   ap %.3e ad %.3e (aff %.3e %.3e) sigma %.3e\n", ap, ad, apA, adA, sigma);
#endif
            }
        }
        if (breakdown && mu <= A.eps_accept && rpmax <= A.eps_p) { status = 0; break; }
        // ================= update =================
        {
            const double a = alpha_step;
            for (int k = 0; k <= NH; ++k) {
                double zk[NZ], dk[NZ], dA[NZ];
                zk[0] = K6W(oSX + 2 * k); zk[1] = K6W(oSX + 2 * k + 1);
                zk[2] = (k > 0) ? K6W(oSX + 2 * (k - 1) + 1) : 0.0;
                dk[0] = K6W(oDSX + 2 * k); dk[1] = K6W(oDSX + 2 * k + 1);
                dk[2] = (k > 0) ? K6W(oDSX + 2 * (k - 1) + 1) : 0.0;
                dA[0] = K6W(oDSXA + 2 * k); dA[1] = K6W(oDSXA + 2 * k + 1);
                dA[2] = (k > 0) ? K6W(oDSXA + 2 * (k - 1) + 1) : 0.0;
                for (int c = 0; c < NB; ++c) {
                    zk[3 + c] = (k > 0) ? K6W(oU + (k - 1) * NU + c) : 0.0;
                    dk[3 + c] = (k > 0) ? K6W(oDU + (k - 1) * NU + c) : 0.0;
                    dA[3 + c] = (k > 0) ? K6W(oDUA + (k - 1) * NU + c) : 0.0;
                }
                for (int c = 0; c < NU; ++c) {
                    zk[NX + c] = (k < NH) ? K6W(oU + k * NU + c) : 0.0;
                    dk[NX + c] = (k < NH) ? K6W(oDU + k * NU + c) : 0.0;
                    dA[NX + c] = (k < NH) ? K6W(oDUA + k * NU + c) : 0.0;
                }
                const int c0 = (k < NH) ? k * rps : NH * rps, c1 = (k < NH) ? c0 + rps : c0 + nterm;
                for (int c = c0; c < c1; ++c) {
                    const double ll = K6W(oLL + c), lu = K6W(oLU + c);
                    if (ll == 0.0 && lu == 0.0) continue;
                    SparseRow r; load_row<NU, NB>(A.rowtab[c], A.rowc, sb, k, r);
                    double az = 0.0, adz = 0.0, adAz = 0.0;
                    for (int e = 0; e < r.nnz; ++e) { az += r.val[e] * zk[r.idx[e]]; adz += r.val[e] * dk[r.idx[e]]; adAz += r.val[e] * dA[r.idx[e]]; }
                    for (int side = 0; side < 2; ++side) {
                        const double lam = side ? lu : ll;
                        if (lam == 0.0) continue;
                        const int ot = side ? oTU : oTL, ol = side ? oLU : oLL;
                        const double t = K6W(ot + c);
                        const double sg = side ? -1.0 : 1.0;
                        const double rp = sg * az - t - (side ? -K6W(oUB + c) : K6W(oLB + c));
                        const double dtA = sg * adAz + rp, dlA = (-t * lam - lam * dtA) / t;
                        const double rc = sigma * mu - t * lam - apA * adA * dtA * dlA;
                        const double dt = sg * adz + rp, dl = (rc - lam * dt) / t;
                        double tn = t + a * dt, ln = lam + a * dl;
                        if (!(tn > 0.0)) tn = 1e-300;           // a == 1 exactly on the boundary: keep the pair interior
                        if (!(ln > 0.0)) ln = 1e-300;
                        K6W(ot + c) = tn; K6W(ol + c) = ln;
                    }
                }
            }
            for (int j = 0; j < n; ++j) K6W(oU + j) += a * K6W(oDU + j);
            for (int j = 0; j < 2 * (NH + 1); ++j) K6W(oSX + j) += a * K6W(oDSX + j);
            nu_res *= (1.0 - a);
        }
    }
    // ---- outputs: controls and the objective 1/2 x'Hx + g'x of the condensed QP (same function of u) ----
    double cost = 0.0;
    for (int k = 0; k <= NH; ++k) {
        const double dv = K6W(oSX + 2 * k + 1);
        cost += K6W(oQZ + k) * dv;
        if (k == NH) break;
        double core[1 + NB], e[1 + NB];
        core[0] = dv; e[0] = (k > 0) ? K6W(oSX + 2 * (k - 1) + 1) : 0.0;
        for (int c = 0; c < NB; ++c) { core[1 + c] = K6W(oU + k * NU + c); e[1 + c] = (k > 0) ? K6W(oU + (k - 1) * NU + c) : 0.0; }
        for (int a = 0; a <= NB; ++a)
            for (int b = 0; b <= NB; ++b) cost += 0.5 * core[a] * sb[EEP_SB_DC + k * 9 + a * 3 + b] * core[b] + core[a] * sb[EEP_SB_CC + k * 9 + a * 3 + b] * e[b];
        for (int c = 0; c < NU; ++c) {
            const double u = K6W(oU + k * NU + c);
            if (c >= NB) cost += 0.5 * sb[EEP_SB_DG + k * 4 + (c - NB)] * u * u;
            cost += K6W(oQZ + NH + 1 + k * NU + c) * u;
            xout[k * NU + c] = (status == 2) ? 0.0 : u;       // numerical breakdown: no command rather than NaN into the plant
        }
    }
    cost_out = (status == 2) ? 0.0 : cost;
    iter_out = it;
    return status;
}

}  // namespace k6
