// proxqp_ref.cpp -- TEST INFRASTRUCTURE (never linked by the product).
//
// Independent cross-check of the QP oracle: the dense ProxQP solver (proxsuite 0.3.2) that the reference
// vendors header-only inside its CasADi 3.6.3 tree (casadi_3_6_3/include/proxsuite, config.hpp:11) is
// compiled from the headers WHERE THEY LIE under /root/reference (oracle/Makefile target `ref`, output
// oracle/_ref/libproxqp_ref.so; no reference source is copied).  qpOASES itself has no source in the
// tree; ProxQP is a different algorithm (augmented-Lagrangian proximal method), so agreement of
// objective values and minimisers pins the oracle's active-set solver at the conic boundary
//     min 1/2 x'Hx + g'x   s.t.  lba <= A x <= uba      (casadi/core/conic.hpp:36-51)
// for the strictly convex ABMPC QPs.  Used only by tests/golden/make_proxqp_golden.py.
#include <proxsuite/proxqp/dense/dense.hpp>
#include <algorithm>

using namespace proxsuite::proxqp;
using Mat = Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::ColMajor>;
using Vec = Eigen::Matrix<double, Eigen::Dynamic, 1>;

// H (n x n) and A (m x n) column-major.  Returns ProxQP's status (0 = solved).
extern "C" int proxqp_ref_solve(int n, int m, const double* H, const double* g, const double* A,
                                const double* lba, const double* uba, double eps_abs, int max_iter,
                                double* x, double* z, double* cost, double* pri_res, double* dua_res, int* iters)
{
    Mat Hm = Eigen::Map<const Mat>(H, n, n);
    Hm = 0.5 * (Hm + Hm.transpose()).eval();
    Vec gv = Eigen::Map<const Vec>(g, n);
    Mat Cm = Eigen::Map<const Mat>(A, m, n);
    Vec l = Eigen::Map<const Vec>(lba, m), u = Eigen::Map<const Vec>(uba, m);
    for (int i = 0; i < m; ++i) { l[i] = std::max(l[i], -1e20); u[i] = std::min(u[i], 1e20); }
    dense::QP<double> qp(n, 0, m);
    qp.settings.eps_abs = eps_abs;
    qp.settings.eps_rel = 0.0;
    qp.settings.max_iter = max_iter;
    qp.settings.initial_guess = InitialGuessStatus::NO_INITIAL_GUESS;    // cold start
    qp.settings.verbose = false;
    qp.init(Hm, gv, proxsuite::nullopt, proxsuite::nullopt, Cm, l, u);
    qp.solve();
    Eigen::Map<Vec>(x, n) = qp.results.x;
    Eigen::Map<Vec>(z, m) = qp.results.z;
    *cost = qp.results.info.objValue;
    *pri_res = qp.results.info.pri_res;
    *dua_res = qp.results.info.dua_res;
    *iters = (int)qp.results.info.iter;
    return (int)qp.results.info.status;
}
