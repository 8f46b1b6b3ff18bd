// Implicit integrators of scipy.integrate.solve_ivp for thread-per-trajectory systems: `method='Radau'` and `method='BDF'`
// (/root/reference/src/pySODM/models/validation.py:158 lists them; base.py:855 hands `method` to solve_ivp with no Jacobian,
// so scipy differentiates the right-hand side numerically).  scipy is a dependency of the reference, not part of it; its
// published algorithms (scipy 1.18, integrate/_ivp/{radau,bdf,common}.py) are restated here STEP FOR STEP -- as for RK45,
// agreement with the reference at its tolerance needs the same step sequence, Newton iterates and Jacobian refreshes:
//   * Radau: 3-stage Radau IIA of order 5 (Hairer & Wanner, Solving ODEs II, IV.8): simplified Newton on the collocation
//     system in the eigenbasis of the Butcher matrix (one real and one complex LU of  mu / h I - J), third-order embedded
//     error estimate filtered by the real LU, predictive (Gustafsson) step-size controller, cubic dense output, the initial
//     Newton guess extrapolated from the previous step's interpolant;
//   * BDF: variable order 1..5 NDF in quasi-constant step form (Shampine & Reichelt, The MATLAB ODE suite; Byrne &
//     Hindmarsh 1975): backward-difference array D, rescaled by R(factor) U on every step-size change, order selection from
//     the error estimates of orders k - 1, k, k + 1 after k + 1 equal steps;
//   * the Jacobian by forward differences with the column-wise adaptive increment of `num_jac` (Salane 1986; Shampine &
//     Reichelt): right-hand-side calls made there are NOT counted in `nfev` (scipy's `fun_vectorized` bypasses the counter);
//   * dense LU with partial pivoting (LAPACK getrf / getrs order: row interchanges, multipliers scaled by the reciprocal
//     pivot, complex pivots chosen by |re| + |im|).
// The steppers expose the interface of `Rk45` (ode_run.cuh): init / step / dense / commit and the fields the output drivers
// read, so both drivers (direct and windowed stores) serve them unchanged.  State lives in per-thread arrays; for systems of at
// most IMPLICIT_STATIC_N equations every index into them is a compile-time constant (see LuReal), larger ones (up to
// IMPLICIT_MAX_N) index at run time and live in local memory.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

namespace sodm {

struct SchemeRadau { static constexpr int NSLOT = 1, WINDOWED_MIN_BLOCKS = 2; };
struct SchemeBDF   { static constexpr int NSLOT = 1, WINDOWED_MIN_BLOCKS = 2; };
constexpr int IMPLICIT_MAX_N = 32;          // dense Jacobian + up to three factorisations per THREAD

__host__ __device__ __forceinline__ bool finite_d(double x) { return fabs(x) <= DBL_MAX; }
// products / sums that must not be contracted into an FMA (the increment of num_jac is defined by its rounding)
__host__ __device__ __forceinline__ double imp_mul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}
__host__ __device__ __forceinline__ double imp_add(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}

// one copy of pow per kernel (its body is several hundred instructions; inlined at every use it crowds the instruction cache)
__host__ __device__ static __noinline__ double imp_pow(double a, double b) { return pow(a, b); }

// ---- dense LU, real and complex ----------------------------------------------------------------------------------------
// Systems of at most IMPLICIT_STATIC_N equations take a path on which EVERY array index is a compile-time constant (loops fully
// unrolled, the pivot row reached through predicated swaps instead of a run-time row index): per-thread arrays indexed at run
// time live in local memory, and at 512 resident threads per SM the 1.5 - 2 KB frames of these kernels overflow L1, so each
// such access is an L2 round trip.  Same arithmetic in the same order on both paths.
constexpr int IMPLICIT_STATIC_N = 8;
#define SODM_UNROLL_SMALL _Pragma("unroll")

template <int N>
struct LuReal {
    double a[N * N];
    int piv[N];
    __host__ __device__ void factor() {
        if constexpr (N <= IMPLICIT_STATIC_N) {
            SODM_UNROLL_SMALL
            for (int k = 0; k < N; ++k) {
                int p = k;
                double best = fabs(a[k * N + k]);
                SODM_UNROLL_SMALL
                for (int i = k + 1; i < N; ++i) {
                    const double v = fabs(a[i * N + k]);
                    if (v > best) { best = v; p = i; }
                }
                piv[k] = p;
                SODM_UNROLL_SMALL
                for (int i = k + 1; i < N; ++i) {
                    const bool sw = p == i;
                    SODM_UNROLL_SMALL
                    for (int j = 0; j < N; ++j) {
                        const double x = a[k * N + j], y = a[i * N + j];
                        a[k * N + j] = sw ? y : x;
                        a[i * N + j] = sw ? x : y;
                    }
                }
                const double d = a[k * N + k];
                const double r = 1.0 / d;
                SODM_UNROLL_SMALL
                for (int i = k + 1; i < N; ++i) {
                    if (d != 0.0) a[i * N + k] *= r;
                    const double l = a[i * N + k];
                    SODM_UNROLL_SMALL
                    for (int j = k + 1; j < N; ++j) a[i * N + j] -= l * a[k * N + j];
                }
            }
        } else {
            for (int k = 0; k < N; ++k) {
                int p = k;
                double best = fabs(a[k * N + k]);
                for (int i = k + 1; i < N; ++i) {
                    const double v = fabs(a[i * N + k]);
                    if (v > best) { best = v; p = i; }
                }
                piv[k] = p;
                if (p != k)
                    for (int j = 0; j < N; ++j) { const double x = a[k * N + j]; a[k * N + j] = a[p * N + j]; a[p * N + j] = x; }
                const double d = a[k * N + k];
                if (d != 0.0) {
                    const double r = 1.0 / d;
                    for (int i = k + 1; i < N; ++i) a[i * N + k] *= r;
                }
                for (int i = k + 1; i < N; ++i) {
                    const double l = a[i * N + k];
                    for (int j = k + 1; j < N; ++j) a[i * N + j] -= l * a[k * N + j];
                }
            }
        }
    }
    __host__ __device__ void solve(double* b) const {
        if constexpr (N <= IMPLICIT_STATIC_N) {
            SODM_UNROLL_SMALL
            for (int k = 0; k < N; ++k) {
                SODM_UNROLL_SMALL
                for (int i = k + 1; i < N; ++i) {
                    const bool sw = piv[k] == i;
                    const double x = b[k], y = b[i];
                    b[k] = sw ? y : x;
                    b[i] = sw ? x : y;
                }
            }
            SODM_UNROLL_SMALL
            for (int i = 1; i < N; ++i) {
                double s = b[i];
                SODM_UNROLL_SMALL
                for (int j = 0; j < i; ++j) s -= a[i * N + j] * b[j];
                b[i] = s;
            }
            SODM_UNROLL_SMALL
            for (int i = N - 1; i >= 0; --i) {
                double s = b[i];
                SODM_UNROLL_SMALL
                for (int j = i + 1; j < N; ++j) s -= a[i * N + j] * b[j];
                b[i] = s / a[i * N + i];
            }
        } else {
            for (int k = 0; k < N; ++k) {
                const int p = piv[k];
                if (p != k) { const double x = b[k]; b[k] = b[p]; b[p] = x; }
            }
            for (int i = 1; i < N; ++i) {
                double s = b[i];
                for (int j = 0; j < i; ++j) s -= a[i * N + j] * b[j];
                b[i] = s;
            }
            for (int i = N - 1; i >= 0; --i) {
                double s = b[i];
                for (int j = i + 1; j < N; ++j) s -= a[i * N + j] * b[j];
                b[i] = s / a[i * N + i];
            }
        }
    }
};

template <int N>
struct LuComplex {
    double re[N * N], im[N * N];
    int piv[N];
    __host__ __device__ void factor() {
        if constexpr (N <= IMPLICIT_STATIC_N) {
            SODM_UNROLL_SMALL
            for (int k = 0; k < N; ++k) {
                int p = k;
                double best = fabs(re[k * N + k]) + fabs(im[k * N + k]);
                SODM_UNROLL_SMALL
                for (int i = k + 1; i < N; ++i) {
                    const double v = fabs(re[i * N + k]) + fabs(im[i * N + k]);
                    if (v > best) { best = v; p = i; }
                }
                piv[k] = p;
                SODM_UNROLL_SMALL
                for (int i = k + 1; i < N; ++i) {
                    const bool sw = p == i;
                    SODM_UNROLL_SMALL
                    for (int j = 0; j < N; ++j) {
                        double x = re[k * N + j], y = re[i * N + j];
                        re[k * N + j] = sw ? y : x;
                        re[i * N + j] = sw ? x : y;
                        x = im[k * N + j]; y = im[i * N + j];
                        im[k * N + j] = sw ? y : x;
                        im[i * N + j] = sw ? x : y;
                    }
                }
                const double dr = re[k * N + k], di = im[k * N + k];
                const bool nz = dr != 0.0 || di != 0.0;
                const double m = 1.0 / (dr * dr + di * di);
                const double rr = dr * m, ri = -di * m;                    // 1 / pivot
                SODM_UNROLL_SMALL
                for (int i = k + 1; i < N; ++i) {
                    if (nz) {
                        const double xr = re[i * N + k], xi = im[i * N + k];
                        re[i * N + k] = xr * rr - xi * ri;
                        im[i * N + k] = xr * ri + xi * rr;
                    }
                    const double lr = re[i * N + k], li = im[i * N + k];
                    SODM_UNROLL_SMALL
                    for (int j = k + 1; j < N; ++j) {
                        const double ur = re[k * N + j], ui = im[k * N + j];
                        re[i * N + j] -= lr * ur - li * ui;
                        im[i * N + j] -= lr * ui + li * ur;
                    }
                }
            }
        } else {
            for (int k = 0; k < N; ++k) {
                int p = k;
                double best = fabs(re[k * N + k]) + fabs(im[k * N + k]);
                for (int i = k + 1; i < N; ++i) {
                    const double v = fabs(re[i * N + k]) + fabs(im[i * N + k]);
                    if (v > best) { best = v; p = i; }
                }
                piv[k] = p;
                if (p != k)
                    for (int j = 0; j < N; ++j) {
                        double x = re[k * N + j]; re[k * N + j] = re[p * N + j]; re[p * N + j] = x;
                        x = im[k * N + j]; im[k * N + j] = im[p * N + j]; im[p * N + j] = x;
                    }
                const double dr = re[k * N + k], di = im[k * N + k];
                if (dr != 0.0 || di != 0.0) {
                    const double m = 1.0 / (dr * dr + di * di);
                    const double rr = dr * m, ri = -di * m;                    // 1 / pivot
                    for (int i = k + 1; i < N; ++i) {
                        const double xr = re[i * N + k], xi = im[i * N + k];
                        re[i * N + k] = xr * rr - xi * ri;
                        im[i * N + k] = xr * ri + xi * rr;
                    }
                }
                for (int i = k + 1; i < N; ++i) {
                    const double lr = re[i * N + k], li = im[i * N + k];
                    for (int j = k + 1; j < N; ++j) {
                        const double ur = re[k * N + j], ui = im[k * N + j];
                        re[i * N + j] -= lr * ur - li * ui;
                        im[i * N + j] -= lr * ui + li * ur;
                    }
                }
            }
        }
    }
    __host__ __device__ void solve(double* br, double* bi) const {
        if constexpr (N <= IMPLICIT_STATIC_N) {
            SODM_UNROLL_SMALL
            for (int k = 0; k < N; ++k) {
                SODM_UNROLL_SMALL
                for (int i = k + 1; i < N; ++i) {
                    const bool sw = piv[k] == i;
                    double x = br[k], y = br[i];
                    br[k] = sw ? y : x;
                    br[i] = sw ? x : y;
                    x = bi[k]; y = bi[i];
                    bi[k] = sw ? y : x;
                    bi[i] = sw ? x : y;
                }
            }
            SODM_UNROLL_SMALL
            for (int i = 1; i < N; ++i) {
                double sr = br[i], si = bi[i];
                SODM_UNROLL_SMALL
                for (int j = 0; j < i; ++j) {
                    sr -= re[i * N + j] * br[j] - im[i * N + j] * bi[j];
                    si -= re[i * N + j] * bi[j] + im[i * N + j] * br[j];
                }
                br[i] = sr; bi[i] = si;
            }
            SODM_UNROLL_SMALL
            for (int i = N - 1; i >= 0; --i) {
                double sr = br[i], si = bi[i];
                SODM_UNROLL_SMALL
                for (int j = i + 1; j < N; ++j) {
                    sr -= re[i * N + j] * br[j] - im[i * N + j] * bi[j];
                    si -= re[i * N + j] * bi[j] + im[i * N + j] * br[j];
                }
                const double dr = re[i * N + i], di = im[i * N + i];
                const double m = 1.0 / (dr * dr + di * di);
                br[i] = (sr * dr + si * di) * m;
                bi[i] = (si * dr - sr * di) * m;
            }
        } else {
            for (int k = 0; k < N; ++k) {
                const int p = piv[k];
                if (p != k) {
                    double x = br[k]; br[k] = br[p]; br[p] = x;
                    x = bi[k]; bi[k] = bi[p]; bi[p] = x;
                }
            }
            for (int i = 1; i < N; ++i) {
                double sr = br[i], si = bi[i];
                for (int j = 0; j < i; ++j) {
                    sr -= re[i * N + j] * br[j] - im[i * N + j] * bi[j];
                    si -= re[i * N + j] * bi[j] + im[i * N + j] * br[j];
                }
                br[i] = sr; bi[i] = si;
            }
            for (int i = N - 1; i >= 0; --i) {
                double sr = br[i], si = bi[i];
                for (int j = i + 1; j < N; ++j) {
                    sr -= re[i * N + j] * br[j] - im[i * N + j] * bi[j];
                    si -= re[i * N + j] * bi[j] + im[i * N + j] * br[j];
                }
                const double dr = re[i * N + i], di = im[i * N + i];
                const double m = 1.0 / (dr * dr + di * di);
                br[i] = (sr * dr + si * di) * m;
                bi[i] = (si * dr - sr * di) * m;
            }
        }
    }
};

// ---- what Radau and BDF share: tolerances, first step, the finite-difference Jacobian ---------------------------------------
template <class Sys>
struct ImplicitBase {
    static constexpr int N = Sys::NL;
    static constexpr int UN = N <= IMPLICIT_STATIC_N ? N : 1;      // unroll factor of the loops over equations
    // read by the output drivers
    double t, t_bound, t_new, h;
    int nfev;
    uint32_t st;
    double ynew[N];
    // solver state
    double rtol, atol, newton_tol, h_abs;
    double J[N * N], jac_factor[N];
    bool have_factor;

    __host__ __device__ static double rms(const double* v, const double* scale) {
        double acc = 0.0;
#pragma unroll UN
        for (int e = 0; e < N; ++e) { const double r = v[e] / scale[e]; acc += r * r; }
        return sqrt(acc) / sqrt((double)N);
    }

    // validate_tol, newton_tol and `select_initial_step` with the method's error-estimator order; f0 = f(t0, y0) is given
    __host__ __device__ void init_common(Sys& sys, const OdeLaunch& L, const double* y, const double* f0, int order) {
        rtol = fmax(L.rtol, 100.0 * DBL_EPSILON);
        atol = L.atol;
        newton_tol = fmax(10.0 * DBL_EPSILON / L.rtol, fmin(0.03, sqrt(L.rtol)));
        t_bound = L.t0 + (double)(L.n_out - 1);
        t = L.t0;
        t_new = t;
        h = 0.0;
        st = ST_OK;
        have_factor = false;
        const double interval = fabs(t_bound - t);
        if (interval == 0.0) { h_abs = 0.0; return; }
        double scale[N], y1[N], f1[N];
#pragma unroll UN
        for (int e = 0; e < N; ++e) scale[e] = atol + fabs(y[e]) * rtol;
        const double d0 = rms(y, scale), d1 = rms(f0, scale);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        h0 = fmin(h0, interval);
#pragma unroll UN
        for (int e = 0; e < N; ++e) y1[e] = y[e] + h0 * f0[e];
        sys.rhs(t + h0, y1, f1); ++nfev;
#pragma unroll UN
        for (int e = 0; e < N; ++e) f1[e] = f1[e] - f0[e];
        const double d2 = rms(f1, scale) / h0;
        const double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : imp_pow(0.01 / fmax(d1, d2), 1.0 / (double)(order + 1));
        h_abs = fmin(fmin(100.0 * h0, h1), interval);
    }

    // `num_jac` (dense): J[:, j] from a forward difference in y_j whose relative increment `jac_factor[j]` persists between
    // calls and is adapted so that the difference stays clear of its own round-off; f = f(tt, yy).  Not counted in nfev.
    __host__ __device__ void num_jac(Sys& sys, double tt, const double* yy, const double* f) {
        const double EPS = DBL_EPSILON;
        // EPS^0.875, EPS^0.75, EPS^0.25 (NUM_JAC_DIFF_REJECT / _SMALL / _BIG)
        const double DIFF_REJECT = 2.0097183471152322e-14, DIFF_SMALL = 1.8189894035458565e-12, DIFF_BIG = 0.0001220703125;
        const double MIN_FACTOR = 1e3 * EPS;
        if (!have_factor) {
            const double f0 = 1.4901161193847656e-08;                          // EPS^0.5
            for (int j = 0; j < N; ++j) jac_factor[j] = f0;
            have_factor = true;
        }
        // the column loop stays a loop (two inlined right-hand sides per column): what it indexes at run time are copies,
        // so that the caller's y and f keep their registers
        double yp[N], yl[N], fl[N], fn[N], diff[N];
#pragma unroll UN
        for (int e = 0; e < N; ++e) { yp[e] = yy[e]; yl[e] = yy[e]; fl[e] = f[e]; }
#pragma unroll 1
        for (int j = 0; j < N; ++j) {
            const double yj = yl[j];
            const double sign = fl[j] >= 0.0 ? 1.0 : -1.0;
            const double y_scale = sign * fmax(atol, fabs(yj));
            double fac = jac_factor[j];
            double hj = imp_add(imp_add(yj, imp_mul(fac, y_scale)), -yj);
            while (hj == 0.0) {
                fac *= 10.0;
                hj = imp_add(imp_add(yj, imp_mul(fac, y_scale)), -yj);
            }
            yp[j] = imp_add(yj, hj);
            sys.rhs(tt, yp, fn);
            double max_diff = -1.0, scale = 0.0;
#pragma unroll UN
            for (int e = 0; e < N; ++e) {
                diff[e] = fn[e] - f[e];
                const double a = fabs(diff[e]);
                if (a > max_diff) { max_diff = a; scale = fmax(fabs(f[e]), fabs(fn[e])); }       // first maximum, as argmax
            }
            if (max_diff < DIFF_REJECT * scale) {
                const double new_fac = 10.0 * fac;
                const double h_new = imp_add(imp_add(yj, imp_mul(new_fac, y_scale)), -yj);
                yp[j] = imp_add(yj, h_new);
                sys.rhs(tt, yp, fn);
                double max_diff_new = -1.0, scale_new = 0.0;
#pragma unroll UN
                for (int e = 0; e < N; ++e) {
                    const double a = fabs(fn[e] - f[e]);
                    if (a > max_diff_new) { max_diff_new = a; scale_new = fmax(fabs(f[e]), fabs(fn[e])); }
                }
                if (max_diff * scale_new < max_diff_new * scale) {
                    fac = new_fac;
                    hj = h_new;
#pragma unroll UN
                    for (int e = 0; e < N; ++e) diff[e] = fn[e] - f[e];
                    scale = scale_new;
                    max_diff = max_diff_new;
                }
            }
            yp[j] = yj;
#pragma unroll UN
            for (int e = 0; e < N; ++e) J[e * N + j] = diff[e] / hj;
            if (max_diff < DIFF_SMALL * scale) fac *= 10.0;
            if (max_diff > DIFF_BIG * scale) fac *= 0.1;
            jac_factor[j] = fmax(fac, MIN_FACTOR);
        }
    }
};

// ---- Radau IIA ------------------------------------------------------------------------------------------------------------
template <class Sys>
struct RadauStepper : ImplicitBase<Sys> {
    using B = ImplicitBase<Sys>;
    static constexpr int N = Sys::NL;
    static constexpr int UN = B::UN;
    static constexpr int NEWTON_MAXITER = 6;
    using B::t; using B::t_bound; using B::t_new; using B::h; using B::nfev; using B::st; using B::ynew;
    using B::rtol; using B::atol; using B::newton_tol; using B::h_abs; using B::J;

    double f[N], fnew[N];
    double h_abs_old, error_norm_old;
    bool have_old, current_jac, lu_valid, have_sol;
    LuReal<N> lur;
    LuComplex<N> luc;
    double Z[3][N];
    double sol_t_old, sol_h, sol_y_old[N], Q[N][3];       // interpolant of the last accepted step

    template <class Stages>
    __host__ __device__ void init(Sys& sys, const OdeLaunch& L, const double* y, Stages&) { init(sys, L, y); }
    __host__ __device__ void init(Sys& sys, const OdeLaunch& L, const double* y) {
        nfev = 0;
        sys.rhs(L.t0, y, f); ++nfev;
        B::init_common(sys, L, y, f, 3);
        have_old = false;
        have_sol = false;
        lu_valid = false;
        B::num_jac(sys, t, y, f);
        current_jac = true;
    }

    // the interpolant of the last accepted step at time tg (also outside its interval: the Newton guess extrapolates)
    __host__ __device__ void sol(double tg, double* out) const {
        const double x = (tg - sol_t_old) / sol_h;
        const double x2 = x * x, x3 = x2 * x;
#pragma unroll UN
        for (int e = 0; e < N; ++e) out[e] = (Q[e][0] * x + Q[e][1] * x2 + Q[e][2] * x3) + sol_y_old[e];
    }

    __host__ __device__ static double predict_factor(double ha, bool old_ok, double ha_old, double en, double en_old) {
        double multiplier = 1.0;
        if (old_ok && en != 0.0) multiplier = ha / ha_old * imp_pow(en_old / en, 0.25);
        return fmin(1.0, multiplier) * imp_pow(en, -0.25);
    }

    // `solve_collocation_system`: simplified Newton iterations for Z (3 x N), started from Z (in/out)
    __host__ __device__ bool newton(Sys& sys, const double* y, const double* scale, int& n_iter, double& rate, bool& rate_ok) {
        // C = [(4 - sqrt 6) / 10, (4 + sqrt 6) / 10, 1]; MU_REAL = 3 + 3^(2/3) - 3^(1/3),
        // MU_COMPLEX = 3 + (3^(1/3) - 3^(2/3)) / 2 - i (3^(5/6) + 3^(7/6)) / 2: the doubles scipy's module-level expressions evaluate to
        const double C[3] = {0.15505102572168222, 0.6449489742783178, 1.0};
        const double MU_REAL = 3.637834252744496, MU_CR = 2.6810828736277523, MU_CI = -3.050430199247411;
        const double T[3][3] = {{0.09443876248897524, -0.14125529502095421, 0.03002919410514742},
                                {0.25021312296533332, 0.20412935229379994, -0.38294211275726192},
                                {1.0, 1.0, 0.0}};
        const double TI[3][3] = {{4.17871859155190428, 0.32768282076106237, 0.52337644549944951},
                                 {-4.17871859155190428, -0.32768282076106237, 0.47662355450055044},
                                 {0.50287263494578682, -2.57192694985560522, 0.59603920482822492}};
        const double M_real = MU_REAL / h, Mcr = MU_CR / h, Mci = MU_CI / h;
        double W[3][N], F[3][N], ytmp[N], br[N], bi[N], b0[N];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll UN
            for (int e = 0; e < N; ++e) W[i][e] = TI[i][0] * Z[0][e] + TI[i][1] * Z[1][e] + TI[i][2] * Z[2][e];
        double dW_norm_old = 0.0;
        bool have_norm = false, converged = false;
        rate_ok = false;
        rate = 0.0;
        int k = 0;
        for (; k < NEWTON_MAXITER; ++k) {
            bool ok = true;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
#pragma unroll UN
                for (int e = 0; e < N; ++e) ytmp[e] = y[e] + Z[i][e];
                sys.rhs(t + h * C[i], ytmp, F[i]); ++nfev;
#pragma unroll UN
                for (int e = 0; e < N; ++e) ok = ok && finite_d(F[i][e]);
            }
            if (!ok) break;
#pragma unroll UN
            for (int e = 0; e < N; ++e) {
                b0[e] = (F[0][e] * TI[0][0] + F[1][e] * TI[0][1] + F[2][e] * TI[0][2]) - M_real * W[0][e];
                // TI_COMPLEX = TI[1] + i TI[2];  M_complex (W1 + i W2)
                const double fr = F[0][e] * TI[1][0] + F[1][e] * TI[1][1] + F[2][e] * TI[1][2];
                const double fi = F[0][e] * TI[2][0] + F[1][e] * TI[2][1] + F[2][e] * TI[2][2];
                br[e] = fr - (Mcr * W[1][e] - Mci * W[2][e]);
                bi[e] = fi - (Mcr * W[2][e] + Mci * W[1][e]);
            }
            lur.solve(b0);
            luc.solve(br, bi);
            double acc = 0.0;
#pragma unroll UN
            for (int e = 0; e < N; ++e) {
                const double r0 = b0[e] / scale[e], r1 = br[e] / scale[e], r2 = bi[e] / scale[e];
                acc += r0 * r0 + r1 * r1 + r2 * r2;
            }
            const double dW_norm = sqrt(acc) / sqrt(3.0 * (double)N);
            if (have_norm) { rate = dW_norm / dW_norm_old; rate_ok = true; }
            if (rate_ok && (rate >= 1.0 || imp_pow(rate, (double)(NEWTON_MAXITER - k)) / (1.0 - rate) * dW_norm > newton_tol)) break;
#pragma unroll UN
            for (int e = 0; e < N; ++e) { W[0][e] += b0[e]; W[1][e] += br[e]; W[2][e] += bi[e]; }
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll UN
                for (int e = 0; e < N; ++e) Z[i][e] = T[i][0] * W[0][e] + T[i][1] * W[1][e] + T[i][2] * W[2][e];
            if (dW_norm == 0.0 || (rate_ok && rate / (1.0 - rate) * dW_norm < newton_tol)) { converged = true; break; }
            dW_norm_old = dW_norm;
            have_norm = true;
        }
        n_iter = k < NEWTON_MAXITER ? k + 1 : NEWTON_MAXITER;
        return converged;
    }

    __host__ __device__ void factorise() {
        const double MU_REAL = 3.637834252744496, MU_CR = 2.6810828736277523, MU_CI = -3.050430199247411;
        const double mr = MU_REAL / h, cr = MU_CR / h, ci = MU_CI / h;
#pragma unroll UN
        for (int i = 0; i < N; ++i)
#pragma unroll UN
            for (int j = 0; j < N; ++j) {
                const double d = i == j ? 1.0 : 0.0;
                lur.a[i * N + j] = mr * d - J[i * N + j];
                luc.re[i * N + j] = cr * d - J[i * N + j];
                luc.im[i * N + j] = ci * d;
            }
        lur.factor();
        luc.factor();
        lu_valid = true;
    }

    template <class Stages>
    __host__ __device__ bool step(Sys& sys, const double* y, Stages&) {
        // C, E = [-13 - 7 sqrt 6, -13 + 7 sqrt 6, -1] / 3 and the interpolation matrix P, as scipy's module evaluates them
        const double C[3] = {0.15505102572168222, 0.6449489742783178, 1.0};
        const double E[3] = {-10.048809399827414, 1.382142733160748, -0.3333333333333333};
        const double P[3][3] = {{10.048809399827414, -25.62959144707664, 15.580782047249224},
                                {-1.382142733160748, 10.296258113743303, -8.914115380582556},
                                {0.3333333333333333, -2.6666666666666665, 3.3333333333333335}};
        const double min_step = 10.0 * (nextafter(t, (double)INFINITY) - t);
        double ha, ha_old = 0.0, en_old = 0.0;
        bool old_ok;
        if (h_abs < min_step) { ha = min_step; old_ok = false; }
        else { ha = h_abs; old_ok = have_old; ha_old = h_abs_old; en_old = error_norm_old; }
        bool rejected = false, rate_ok = false;
        int n_iter = 0;
        double rate = 0.0, error_norm = 0.0, safety = 0.0;
        double scale[N], err[N], ze[N], ytmp[N];
        for (;;) {
            if (ha < min_step) { st |= ST_STEP_TOO_SMALL; return false; }
            h = ha;
            t_new = t + h;
            if (t_new - t_bound > 0.0) t_new = t_bound;
            h = t_new - t;
            ha = fabs(h);
#pragma unroll UN
            for (int e = 0; e < N; ++e) scale[e] = atol + fabs(y[e]) * rtol;
            bool converged = false;
            for (;;) {
                if (!lu_valid) factorise();
                if (!have_sol) {
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll UN
                        for (int e = 0; e < N; ++e) Z[i][e] = 0.0;
                } else {
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        sol(t + h * C[i], ytmp);
#pragma unroll UN
                        for (int e = 0; e < N; ++e) Z[i][e] = ytmp[e] - y[e];
                    }
                }
                converged = newton(sys, y, scale, n_iter, rate, rate_ok);
                if (converged || current_jac) break;
                B::num_jac(sys, t, y, f);
                current_jac = true;
                lu_valid = false;
            }
            if (!converged) {
                ha *= 0.5;
                lu_valid = false;
                continue;
            }
#pragma unroll UN
            for (int e = 0; e < N; ++e) {
                ynew[e] = y[e] + Z[2][e];
                ze[e] = (Z[0][e] * E[0] + Z[1][e] * E[1] + Z[2][e] * E[2]) / h;
                err[e] = f[e] + ze[e];
            }
            lur.solve(err);
#pragma unroll UN
            for (int e = 0; e < N; ++e) scale[e] = atol + fmax(fabs(y[e]), fabs(ynew[e])) * rtol;
            error_norm = B::rms(err, scale);
            safety = 0.9 * (2 * NEWTON_MAXITER + 1) / (double)(2 * NEWTON_MAXITER + n_iter);
            if (rejected && error_norm > 1.0) {
#pragma unroll UN
                for (int e = 0; e < N; ++e) ytmp[e] = y[e] + err[e];
                sys.rhs(t, ytmp, err); ++nfev;
#pragma unroll UN
                for (int e = 0; e < N; ++e) err[e] = err[e] + ze[e];
                lur.solve(err);
                error_norm = B::rms(err, scale);
            }
            if (!(error_norm == error_norm)) { st |= ST_NONFINITE | ST_STEP_TOO_SMALL; return false; }
            if (error_norm > 1.0) {
                const double factor = predict_factor(ha, old_ok, ha_old, error_norm, en_old);
                ha *= fmax(0.2, safety * factor);
                lu_valid = false;
                rejected = true;
            } else break;
        }
        const bool recompute_jac = n_iter > 2 && rate > 1e-3;
        double factor = predict_factor(ha, old_ok, ha_old, error_norm, en_old);
        factor = fmin(10.0, safety * factor);
        if (!recompute_jac && factor < 1.2) factor = 1.0;
        else lu_valid = false;
        sys.rhs(t_new, ynew, fnew); ++nfev;
        if (recompute_jac) { B::num_jac(sys, t_new, ynew, fnew); current_jac = true; }
        else current_jac = false;
        h_abs_old = h_abs;
        error_norm_old = error_norm;
        have_old = true;
        h_abs = ha * factor;
        sol_t_old = t;
        sol_h = t_new - t;
#pragma unroll UN
        for (int e = 0; e < N; ++e) {
            sol_y_old[e] = y[e];
            for (int j = 0; j < 3; ++j) Q[e][j] = Z[0][e] * P[0][j] + Z[1][e] * P[1][j] + Z[2][e] * P[2][j];
        }
        have_sol = true;
        return true;
    }

    template <class Stages>
    __host__ __device__ void dense(Sys&, const double*, Stages&, double tg, double* out) { sol(tg, out); }

    __host__ __device__ void commit(double* y) {
#pragma unroll UN
        for (int e = 0; e < N; ++e) { y[e] = ynew[e]; f[e] = fnew[e]; }
        t = t_new;
    }
};

// ---- BDF / NDF ------------------------------------------------------------------------------------------------------------
template <class Sys>
struct BdfStepper : ImplicitBase<Sys> {
    using B = ImplicitBase<Sys>;
    static constexpr int N = Sys::NL;
    static constexpr int UN = B::UN;
    static constexpr int MAX_ORDER = 5, NEWTON_MAXITER = 4;
    using B::t; using B::t_bound; using B::t_new; using B::h; using B::nfev; using B::st; using B::ynew;
    using B::rtol; using B::atol; using B::newton_tol; using B::h_abs; using B::J;

    double D[MAX_ORDER + 3][N];               // indexed by compile-time constants only (unrolled loops, predicated on the order)
    int order, n_equal_steps;
    // kappa = [0, -0.1850, -1/9, -0.0823, -0.0415, 0]; gamma = [0, cumsum 1/k]; alpha = (1 - kappa) gamma;
    // error_const = kappa gamma + 1 / (k + 1): the doubles scipy's expressions evaluate to, selected without a table in memory
    __host__ __device__ static __forceinline__ double pick6(int k, double c0, double c1, double c2, double c3, double c4, double c5) {
        return k == 0 ? c0 : k == 1 ? c1 : k == 2 ? c2 : k == 3 ? c3 : k == 4 ? c4 : c5;
    }
    __host__ __device__ static __forceinline__ double gamma_(int k) { return pick6(k, 0.0, 1.0, 1.5, 1.8333333333333333, 2.083333333333333, 2.283333333333333); }
    __host__ __device__ static __forceinline__ double alpha_(int k) {
        return pick6(k, 0.0, 1.185, 1.6666666666666667, 1.9842166666666667, 2.1697916666666663, 2.283333333333333);
    }
    __host__ __device__ static __forceinline__ double error_const_(int k) {
        return pick6(k, 1.0, 0.315, 0.16666666666666666, 0.09911666666666669, 0.11354166666666668, 0.16666666666666666);
    }
    bool lu_valid;
    LuReal<N> lu;

    // `change_D`: D[:order + 1] <- (R(factor) U)^T D[:order + 1].  R[i][j] = prod_{m <= i} (m - 1 - factor j) / m does not depend
    // on the order, and U = R(1) is the signed Pascal matrix U[k][j] = (-1)^k C(j, k) exactly in floating point (checked against
    // scipy's compute_R), so R U is formed for the maximal order with every loop unrolled -- its leading (order + 1)^2 block is
    // what scipy computes, the same terms in the same order -- and applied under predicates on the order.
    __host__ __device__ void change_D(int ord, double factor) {
        constexpr int M = MAX_ORDER + 1;
        constexpr double U[M][M] = {{1, 1, 1, 1, 1, 1}, {0, -1, -2, -3, -4, -5}, {0, 0, 1, 3, 6, 10},
                                    {0, 0, 0, -1, -4, -10}, {0, 0, 0, 0, 1, 5}, {0, 0, 0, 0, 0, -1}};
        double R[M][M], RU[M][M];
#pragma unroll
        for (int j = 0; j < M; ++j) R[0][j] = 1.0;
#pragma unroll
        for (int i = 1; i < M; ++i) {
            R[i][0] = 0.0;
#pragma unroll
            for (int j = 1; j < M; ++j) R[i][j] = R[i - 1][j] * (((double)(i - 1) - factor * (double)j) / (double)i);
        }
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int j = 0; j < M; ++j) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k <= j; ++k) s += R[i][k] * U[k][j];           // U[k][j] = 0 for k > j
                RU[i][j] = s;
            }
#pragma unroll UN
        for (int e = 0; e < N; ++e) {
            double col[M];
#pragma unroll
            for (int i = 0; i < M; ++i) col[i] = D[i][e];
#pragma unroll
            for (int i = 0; i < M; ++i) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < M; ++k) s += k <= ord ? RU[k][i] * col[k] : 0.0;
                if (i <= ord) D[i][e] = s;
            }
        }
    }

    template <class Stages>
    __host__ __device__ void init(Sys& sys, const OdeLaunch& L, const double* y, Stages&) { init(sys, L, y); }
    __host__ __device__ void init(Sys& sys, const OdeLaunch& L, const double* y) {
        double f[N];
        nfev = 0;
        sys.rhs(L.t0, y, f); ++nfev;
        B::init_common(sys, L, y, f, 1);
        B::num_jac(sys, t, y, f);
#pragma unroll UN
        for (int e = 0; e < N; ++e) { D[0][e] = y[e]; D[1][e] = f[e] * h_abs; }
#pragma unroll
        for (int k = 2; k < MAX_ORDER + 3; ++k)
#pragma unroll UN
            for (int e = 0; e < N; ++e) D[k][e] = 0.0;
        order = 1;
        n_equal_steps = 0;
        lu_valid = false;
    }

    __host__ __device__ bool newton(Sys& sys, const double* y_predict, double c, const double* psi, const double* scale,
                                    int& n_iter, double* d) {
        double f[N], dy[N];
#pragma unroll UN
        for (int e = 0; e < N; ++e) { d[e] = 0.0; ynew[e] = y_predict[e]; }
        double dy_norm_old = 0.0;
        bool have_norm = false, converged = false;
        int k = 0;
        for (; k < NEWTON_MAXITER; ++k) {
            sys.rhs(t_new, ynew, f); ++nfev;
            bool ok = true;
#pragma unroll UN
            for (int e = 0; e < N; ++e) ok = ok && finite_d(f[e]);
            if (!ok) break;
#pragma unroll UN
            for (int e = 0; e < N; ++e) dy[e] = (c * f[e] - psi[e]) - d[e];
            lu.solve(dy);
            const double dy_norm = B::rms(dy, scale);
            double rate = 0.0;
            const bool rate_ok = have_norm;
            if (rate_ok) rate = dy_norm / dy_norm_old;
            if (rate_ok && (rate >= 1.0 || imp_pow(rate, (double)(NEWTON_MAXITER - k)) / (1.0 - rate) * dy_norm > newton_tol)) break;
#pragma unroll UN
            for (int e = 0; e < N; ++e) { ynew[e] += dy[e]; d[e] += dy[e]; }
            if (dy_norm == 0.0 || (rate_ok && rate / (1.0 - rate) * dy_norm < newton_tol)) { converged = true; break; }
            dy_norm_old = dy_norm;
            have_norm = true;
        }
        n_iter = k < NEWTON_MAXITER ? k + 1 : NEWTON_MAXITER;
        return converged;
    }

    template <class Stages>
    __host__ __device__ bool step(Sys& sys, const double*, Stages&) {
        const double min_step = 10.0 * (nextafter(t, (double)INFINITY) - t);
        double ha;
        // `change_D` has ONE call site: every rescaling of the difference array is posted as a pending factor and applied at the
        // top of the attempt loop (inlined at each of scipy's six call sites the kernel was 258 KB of code and spent 58 % of its
        // stall cycles waiting for instructions, profiles/r2_sir_bdf_ncu_summary.txt).  Same calls, same order, same arguments.
        double pend = 0.0;                       // factor of the posted rescaling (0: none)
        if (h_abs < min_step) {
            ha = min_step;
            pend = min_step / h_abs;
            n_equal_steps = 0;
        } else ha = h_abs;
        bool current_jac = false, clipped = false, accepted = false;
        double y_predict[N], scale[N], psi[N], d[N], tmp[N];
        double error_norm = 0.0, safety = 0.0;
        int n_iter = 0;
        for (;;) {
            if (pend != 0.0) { change_D(order, pend); pend = 0.0; }
            if (accepted) return true;
            if (!clipped) {
                if (ha < min_step) { st |= ST_STEP_TOO_SMALL; return false; }
                h = ha;
                t_new = t + h;
                if (t_new - t_bound > 0.0) {
                    t_new = t_bound;
                    pend = fabs(t_new - t) / ha;
                    n_equal_steps = 0;
                    lu_valid = false;
                    clipped = true;
                }
                h = t_new - t;
                ha = fabs(h);
                if (clipped) continue;
            }
            clipped = false;
            const double alpha_o = alpha_(order);
#pragma unroll UN
            for (int e = 0; e < N; ++e) {
                double s = D[0][e];
#pragma unroll
                for (int k = 1; k <= MAX_ORDER; ++k) s += k <= order ? D[k][e] : 0.0;
                y_predict[e] = s;
                scale[e] = atol + rtol * fabs(s);
                double p = 0.0;
#pragma unroll
                for (int k = 1; k <= MAX_ORDER; ++k) p += k <= order ? D[k][e] * gamma_(k) : 0.0;
                psi[e] = p / alpha_o;
            }
            const double c = h / alpha_o;
            bool converged = false;
            for (;;) {
                if (!lu_valid) {
#pragma unroll UN
                    for (int i = 0; i < N; ++i)
#pragma unroll UN
                        for (int j = 0; j < N; ++j) lu.a[i * N + j] = (i == j ? 1.0 : 0.0) - c * J[i * N + j];
                    lu.factor();
                    lu_valid = true;
                }
                converged = newton(sys, y_predict, c, psi, scale, n_iter, d);
                if (converged || current_jac) break;
                sys.rhs(t_new, y_predict, tmp);                     // `fun_single`: not counted
                B::num_jac(sys, t_new, y_predict, tmp);
                lu_valid = false;
                current_jac = true;
            }
            if (!converged) {
                ha *= 0.5;
                pend = 0.5;
                n_equal_steps = 0;
                lu_valid = false;
                continue;
            }
            safety = 0.9 * (2 * NEWTON_MAXITER + 1) / (double)(2 * NEWTON_MAXITER + n_iter);
#pragma unroll UN
            for (int e = 0; e < N; ++e) {
                scale[e] = atol + rtol * fabs(ynew[e]);
                tmp[e] = error_const_(order) * d[e];
            }
            error_norm = B::rms(tmp, scale);
            if (!(error_norm == error_norm)) { st |= ST_NONFINITE | ST_STEP_TOO_SMALL; return false; }
            if (error_norm > 1.0) {
                const double factor = fmax(0.2, safety * imp_pow(error_norm, -1.0 / (double)(order + 1)));
                ha *= factor;
                pend = factor;
                n_equal_steps = 0;
                // (scipy keeps the factorisation here: c changes with h, but LU is only reset where the code above says so)
                continue;
            }
            // ---- accepted ------------------------------------------------------------------------------------------------
            n_equal_steps += 1;
            h_abs = ha;
#pragma unroll UN
            for (int e = 0; e < N; ++e) {
#pragma unroll
                for (int k = 2; k <= MAX_ORDER + 2; ++k) {
                    if (k == order + 2) { D[k][e] = d[e] - D[k - 1][e]; D[k - 1][e] = d[e]; }
                }
#pragma unroll
                for (int i = MAX_ORDER; i >= 0; --i) {
                    if (i <= order) D[i][e] += D[i + 1][e];
                }
            }
            if (n_equal_steps < order + 1) return true;
            double en[3];
            en[0] = INFINITY;
            en[1] = error_norm;
            en[2] = INFINITY;
#pragma unroll 1
            for (int side = 0; side < 3; side += 2) {               // the estimates of orders k - 1 (D[k]) and k + 1 (D[k + 2])
                if (side == 0 ? order > 1 : order < MAX_ORDER) {
                    const double ec = error_const_(order - 1 + side);
#pragma unroll UN
                    for (int e = 0; e < N; ++e) tmp[e] = ec * d_at(order + side, e);
                    en[side] = B::rms(tmp, scale);
                }
            }
            int best = 0;
            double fmaxv = -1.0;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double fi = imp_pow(en[i], -1.0 / (double)(order + i));
                if (fi > fmaxv) { fmaxv = fi; best = i; }
            }
            order += best - 1;
            const double factor = fmin(10.0, safety * fmaxv);
            h_abs *= factor;
            pend = factor;
            n_equal_steps = 0;
            lu_valid = false;
            accepted = true;
        }
    }

    // D[k][e] for a run-time k in 1 .. MAX_ORDER + 2
    __host__ __device__ __forceinline__ double d_at(int k, int e) const {
        double v = 0.0;
#pragma unroll
        for (int q = 1; q <= MAX_ORDER + 2; ++q) v = q == k ? D[q][e] : v;
        return v;
    }

    // BdfDenseOutput, built from the state AFTER the step (new order, new step size, rescaled differences)
    template <class Stages>
    __host__ __device__ void dense(Sys&, const double*, Stages&, double tg, double* out) {
        double p[MAX_ORDER];
        double acc = 1.0;
#pragma unroll
        for (int k = 0; k < MAX_ORDER; ++k) {
            p[k] = 0.0;
            if (k < order) {
                const double x = (tg - (t_new - h_abs * (double)k)) / (h_abs * (double)(1 + k));
                acc *= x;
                p[k] = acc;
            }
        }
#pragma unroll UN
        for (int e = 0; e < N; ++e) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < MAX_ORDER; ++k) s += k < order ? D[1 + k][e] * p[k] : 0.0;
            out[e] = s + D[0][e];
        }
    }

    __host__ __device__ void commit(double* y) {
#pragma unroll UN
        for (int e = 0; e < N; ++e) y[e] = ynew[e];
        t = t_new;
    }
};

}  // namespace sodm
