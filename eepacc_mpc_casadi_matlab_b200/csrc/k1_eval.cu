// k1_eval.cu -- K1: batched evaluator of the reference's CasADi expressions, one point per thread.
//
//   f      : (x[4]=[s v theta j], u[6]=[Fm Fb xi_v xi_h xi_s xi_f]) -> (xdot[4], L)   MC/RunOpt_NLP.m:225-245,263
//   F      : 4 RK4 steps of f over Ts -> (xf[4], qf)                                  MC/RunOpt_NLP.m:275-288
//   F_jac  : d[xf;qf]/d[x0;p] (5x10, column-major) by forward-mode dual numbers
//   power5 / power2 : F/Other/GetMotorPower_{FifthOrderSurface,QuadraticFit}.m
//   plant  : F/MPCs/RunPlantModel.m (K4, 10 RK4 sub-steps with PWA slope lookup)
//
// Layout: batch-leading AoS exactly as the C-ABI hands it over (double[B][nnz]); each thread
// moves its 32/48-byte records with 16-byte vector loads, adjacent threads touch adjacent
// records, so every 128-byte line is fully used.  HBM bytes per point: f 80+40, F 80+40,
// F_jac 80+440.
#include "eepacc_dev.cuh"

template <int ND>
struct Dual {
    double v;
    double d[ND];
};
template <int ND> __device__ __forceinline__ Dual<ND> dconst(double c) { Dual<ND> r; r.v = c;
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = 0.0; return r; }
template <int ND> __device__ __forceinline__ Dual<ND> operator+(Dual<ND> a, const Dual<ND>& b) { a.v += b.v;
#pragma unroll
    for (int i = 0; i < ND; ++i) a.d[i] += b.d[i]; return a; }
template <int ND> __device__ __forceinline__ Dual<ND> operator-(Dual<ND> a, const Dual<ND>& b) { a.v -= b.v;
#pragma unroll
    for (int i = 0; i < ND; ++i) a.d[i] -= b.d[i]; return a; }
template <int ND> __device__ __forceinline__ Dual<ND> operator*(const Dual<ND>& a, const Dual<ND>& b) { Dual<ND> r; r.v = a.v * b.v;
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int ND> __device__ __forceinline__ Dual<ND> operator*(double c, Dual<ND> a) { a.v *= c;
#pragma unroll
    for (int i = 0; i < ND; ++i) a.d[i] *= c; return a; }
template <int ND> __device__ __forceinline__ Dual<ND> dcos(const Dual<ND>& a) { Dual<ND> r; double s, c; sincos(a.v, &s, &c); r.v = c;
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = -s * a.d[i]; return r; }
template <int ND> __device__ __forceinline__ Dual<ND> dsin(const Dual<ND>& a) { Dual<ND> r; double s, c; sincos(a.v, &s, &c); r.v = s;
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = c * a.d[i]; return r; }

// plain-double overloads so the same template body serves ND = 0
struct D0 { double v; };
__device__ __forceinline__ D0 operator+(D0 a, const D0& b) { a.v += b.v; return a; }
__device__ __forceinline__ D0 operator-(D0 a, const D0& b) { a.v -= b.v; return a; }
__device__ __forceinline__ D0 operator*(const D0& a, const D0& b) { D0 r; r.v = a.v * b.v; return r; }
__device__ __forceinline__ D0 operator*(double c, D0 a) { a.v *= c; return a; }
__device__ __forceinline__ D0 dcos(const D0& a) { D0 r; r.v = cos(a.v); return r; }
__device__ __forceinline__ D0 dsin(const D0& a) { D0 r; r.v = sin(a.v); return r; }

// stage cost + dynamics (RunOpt_NLP.m:225-245).  theta, j are constant states (xdot = 0).
template <class T>
__device__ __forceinline__ void eval_f(const eepacc_params& p, double m, const T& v, const T& th, const T& j,
                                       const T* u, T& acc, T& L)
{
    const double* W = p.W_NLP;
    const T& Fm = u[0];
    const T& Fb = u[1];
    acc = Fm + Fb;
    acc = acc - p.zeta_a * (v * v);
    acc = acc - (p.c_r * m * p.g) * dcos(th);
    acc = acc - (m * p.g) * dsin(th);
    acc = (1.0 / (p.lambda * m)) * acc;
    T rpm = ((30.0 / EEP_PI) * p.phi) * v;
    T P;
    if (p.nlp_use_fifth) {
        const double* b = p.b_fifth;
        T x2 = Fm * Fm, x3 = x2 * Fm, x4 = x3 * Fm, x5 = x4 * Fm;
        T y2 = rpm * rpm, y3 = y2 * rpm, y4 = y3 * rpm, y5 = y4 * rpm;
        P = b[1] * Fm;  P.v += b[0];
        P = P + b[2] * rpm;          P = P + b[3] * x2;           P = P + b[4] * (Fm * rpm);
        P = P + b[5] * y2;           P = P + b[6] * x3;           P = P + b[7] * (x2 * rpm);
        P = P + b[8] * (Fm * y2);    P = P + b[9] * y3;           P = P + b[10] * x4;
        P = P + b[11] * (x3 * rpm);  P = P + b[12] * (x2 * y2);   P = P + b[13] * (Fm * y3);
        P = P + b[14] * y4;          P = P + b[15] * x5;          P = P + b[16] * (x4 * rpm);
        P = P + b[17] * (x3 * y2);   P = P + b[18] * (x2 * y3);   P = P + b[19] * (Fm * y4);
        P = P + b[20] * y5;
    } else {
        const double* b = p.b_quadr;
        P = b[1] * Fm;  P.v += b[0];
        P = P + b[2] * rpm; P = P + b[3] * (Fm * Fm); P = P + b[4] * (Fm * rpm); P = P + b[5] * (rpm * rpm);
    }
    L = W[0] * P;
    L = L + W[1] * (acc * acc);
    L = L + W[2] * (j * j);
    L = L + W[3] * u[2];
    L = L + W[4] * (u[3] * u[3] + 1e2 * u[3]);
    L = L + W[5] * u[4];
    L = L + W[6] * u[5];
}

// RunOpt_NLP.m:275-288, M = 4.  Only s and v move; xdot(3:4) = 0.
template <class T>
__device__ __forceinline__ void eval_F(const eepacc_params& p, double m, T& s, T& v, const T& th, const T& j,
                                       const T* u, T& Q)
{
    const int M = 4;
    const double DT = p.Ts / M;
    bool first = true;
    for (int ii = 0; ii < M; ++ii) {
        T a1, a2, a3, a4, q1, q2, q3, q4;
        eval_f(p, m, v, th, j, u, a1, q1);
        T v2 = v + (DT / 2) * a1;
        eval_f(p, m, v2, th, j, u, a2, q2);
        T v3 = v + (DT / 2) * a2;
        eval_f(p, m, v3, th, j, u, a3, q3);
        T v4 = v + DT * a3;
        eval_f(p, m, v4, th, j, u, a4, q4);
        // k_s = the velocity argument of each stage
        T ks = (v + 2.0 * v2) + (2.0 * v3 + v4);
        T kv = (a1 + 2.0 * a2) + (2.0 * a3 + a4);
        T kq = (q1 + 2.0 * q2) + (2.0 * q3 + q4);
        s = s + (DT / 6) * ks;
        v = v + (DT / 6) * kv;
        if (first) { Q = (DT / 6) * kq; first = false; }
        else Q = Q + (DT / 6) * kq;
    }
}

__device__ __forceinline__ void load4(const double* __restrict__ p, double* o)
{
    const double2* q = reinterpret_cast<const double2*>(p);
    double2 a = q[0], b = q[1];
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
__device__ __forceinline__ void load6(const double* __restrict__ p, double* o)
{
    const double2* q = reinterpret_cast<const double2*>(p);
    double2 a = q[0], b = q[1], c = q[2];
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y; o[4] = c.x; o[5] = c.y;
}

__global__ void __launch_bounds__(256) k1_f_kernel(eepacc_params p, long long B, const double* __restrict__ x,
                                                   const double* __restrict__ u, double* __restrict__ xdot,
                                                   double* __restrict__ L)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < B; i += stride) {
        double xv[4], uv[6];
        load4(x + 4 * i, xv); load6(u + 6 * i, uv);
        D0 U[6]; for (int k = 0; k < 6; ++k) U[k].v = uv[k];
        D0 v{xv[1]}, th{xv[2]}, j{xv[3]}, acc, Lv;
        eval_f(p, p.m, v, th, j, U, acc, Lv);
        double2* o = reinterpret_cast<double2*>(xdot + 4 * i);
        o[0] = make_double2(xv[1], acc.v);
        o[1] = make_double2(0.0, 0.0);
        L[i] = Lv.v;
    }
}

__global__ void __launch_bounds__(256) k1_F_kernel(eepacc_params p, long long B, const double* __restrict__ x,
                                                   const double* __restrict__ u, double* __restrict__ xf,
                                                   double* __restrict__ qf)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < B; i += stride) {
        double xv[4], uv[6];
        load4(x + 4 * i, xv); load6(u + 6 * i, uv);
        D0 U[6]; for (int k = 0; k < 6; ++k) U[k].v = uv[k];
        D0 s{xv[0]}, v{xv[1]}, th{xv[2]}, j{xv[3]}, Q;
        eval_F(p, p.m, s, v, th, j, U, Q);
        double2* o = reinterpret_cast<double2*>(xf + 4 * i);
        o[0] = make_double2(s.v, v.v);
        o[1] = make_double2(xv[2], xv[3]);
        qf[i] = Q.v;
    }
}

// Active derivative directions: [v0, theta, j, Fm, Fb, xi_h].  s0 only shifts xf(1) (d xf1/d s0 = 1),
// xi_v / xi_s / xi_f enter L linearly with constant weights, so their columns are the RK4
// quadrature of a constant: exactly Ts * w (RunOpt_NLP.m:245).  xi_h is kept as a dual
// because of the xi_h^2 term.
__global__ void __launch_bounds__(128) k1_Fjac_kernel(eepacc_params p, long long B, const double* __restrict__ x,
                                                      const double* __restrict__ u, double* __restrict__ xf,
                                                      double* __restrict__ qf, double* __restrict__ J)
{
    constexpr int ND = 6;
    typedef Dual<ND> T;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < B; i += stride) {
        double xv[4], uv[6];
        load4(x + 4 * i, xv); load6(u + 6 * i, uv);
        T U[6];
        for (int k = 0; k < 6; ++k) U[k] = dconst<ND>(uv[k]);
        U[0].d[3] = 1.0; U[1].d[4] = 1.0; U[3].d[5] = 1.0;
        T s = dconst<ND>(xv[0]), v = dconst<ND>(xv[1]), th = dconst<ND>(xv[2]), j = dconst<ND>(xv[3]), Q;
        v.d[0] = 1.0; th.d[1] = 1.0; j.d[2] = 1.0;
        eval_F(p, p.m, s, v, th, j, U, Q);
        double2* o = reinterpret_cast<double2*>(xf + 4 * i);
        o[0] = make_double2(s.v, v.v);
        o[1] = make_double2(xv[2], xv[3]);
        qf[i] = Q.v;
        // J (5 x 10, column-major): columns [s0 v0 theta j Fm Fb xi_v xi_h xi_s xi_f]
        double* Jp = J + 50 * i;
        const int map[10] = { -1, 0, 1, 2, 3, 4, -2, 5, -3, -4 };
        // RK4 quadrature weight of a constant integrand, accumulated like eval_F does
        const double DT = p.Ts / 4;
        double qc = 0.0;
        for (int ii = 0; ii < 4; ++ii) { double t = (DT / 6) * ((1.0 + 2.0) + (2.0 + 1.0)); qc = (ii == 0) ? t : qc + t; }
#pragma unroll
        for (int c = 0; c < 10; ++c) {
            double col[5];
            int k = map[c];
            if (k >= 0) {
                col[0] = s.d[k]; col[1] = v.d[k]; col[2] = (c == 2) ? 1.0 : 0.0; col[3] = (c == 3) ? 1.0 : 0.0;
                col[4] = Q.d[k];
            } else {
                col[0] = (k == -1) ? 1.0 : 0.0; col[1] = 0.0; col[2] = 0.0; col[3] = 0.0;
                col[4] = (k == -2) ? qc * p.W_NLP[3] : (k == -3) ? qc * p.W_NLP[5] : (k == -4) ? qc * p.W_NLP[6] : 0.0;
            }
#pragma unroll
            for (int r = 0; r < 5; ++r) Jp[c * 5 + r] = col[r];
        }
    }
}

__global__ void __launch_bounds__(256) k1_power_kernel(eepacc_params p, long long B, int fifth,
                                                       const double* __restrict__ Fm, const double* __restrict__ rpm,
                                                       double* __restrict__ P)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < B; i += stride) P[i] = fifth ? eep_power5(p.b_fifth, Fm[i], rpm[i]) : eep_power2(p.b_quadr, Fm[i], rpm[i]);
}

// K4: (x[2], u[2]) -> x_next[2]; the route block of vehicle i supplies the slope PWA.
// routes == nullptr-safe: a single shared route (stride 0) is allowed.
__global__ void __launch_bounds__(256) k4_plant_kernel(eepacc_params p, long long B, const eepacc_route* __restrict__ routes,
                                                       long long route_stride, const double* __restrict__ x,
                                                       const double* __restrict__ u, double* __restrict__ xn)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < B; i += stride) {
        const double2 xv = reinterpret_cast<const double2*>(x)[i];
        const double2 uv = reinterpret_cast<const double2*>(u)[i];
        double s = xv.x, v = xv.y;
        eep_plant_step(p, routes + i * route_stride, s, v, uv.x, uv.y);
        reinterpret_cast<double2*>(xn)[i] = make_double2(s, v);
    }
}

static inline int grid_for(long long B, int bs)
{
    long long g = (B + bs - 1) / bs;
    const long long cap = 148LL * 16;          // 148 SMs x resident CTAs; grid-stride beyond that
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

extern "C" {
void eep_launch_f(const eepacc_params* p, long long B, const double* x, const double* u, double* xdot, double* L, cudaStream_t st)
{ k1_f_kernel<<<grid_for(B, 256), 256, 0, st>>>(*p, B, x, u, xdot, L); }
void eep_launch_F(const eepacc_params* p, long long B, const double* x, const double* u, double* xf, double* qf, cudaStream_t st)
{ k1_F_kernel<<<grid_for(B, 256), 256, 0, st>>>(*p, B, x, u, xf, qf); }
void eep_launch_Fjac(const eepacc_params* p, long long B, const double* x, const double* u, double* xf, double* qf, double* J, cudaStream_t st)
{ k1_Fjac_kernel<<<grid_for(B, 128), 128, 0, st>>>(*p, B, x, u, xf, qf, J); }
void eep_launch_power(const eepacc_params* p, long long B, int fifth, const double* Fm, const double* rpm, double* P, cudaStream_t st)
{ k1_power_kernel<<<grid_for(B, 256), 256, 0, st>>>(*p, B, fifth, Fm, rpm, P); }
void eep_launch_plant(const eepacc_params* p, long long B, const eepacc_route* routes, long long route_stride,
                      const double* x, const double* u, double* xn, cudaStream_t st)
{ k4_plant_kernel<<<grid_for(B, 256), 256, 0, st>>>(*p, B, routes, route_stride, x, u, xn); }
}
