// Per-trajectory ODE integrators.
//
// `rk45_run` follows scipy.integrate.solve_ivp(method='RK45') STEP FOR STEP, because that is the call the
// reference makes (/root/reference/src/pySODM/models/base.py:855) and the parity bar (rtol 1e-6 against
// it at the default rtol=1e-4) can only be met by reproducing its step sequence.  The
// published algorithm restated here (scipy 1.18, integrate/_ivp/{rk,common,ivp,base}.py):
//   * Dormand-Prince 5(4) pair with FSAL, local extrapolation (the 5th-order solution is propagated);
//   * error norm  = RMS( h * K^T E / (atol + max(|y|,|y_new|) * rtol) );
//   * accept if norm < 1, next h_abs *= min(10, 0.9 * norm^(-1/5)) (capped at 1 after a rejection in the
//     same step; 10 if norm == 0); reject: h_abs *= max(0.2, 0.9 * norm^(-1/5));
//   * h_abs clamped from below by min_step = 10 * ulp(t); the last step is clipped to t_bound;
//   * first step from Hairer's heuristic (`select_initial_step`, order 4);
//   * output at `t_eval` from the quartic dense-output polynomial of the step that covers it
//     (Shampine 1986; `RkDenseOutput`: y_old + h * Q p,  Q = K^T P,  p = [x, x^2, x^3, x^4]).
// `euler_run` follows `ODE._solve_discrete` (base.py:800-819): explicit Euler + np.interp onto the grid.
// `rk4_run` is an extension (classical RK4, same output rule); the reference has no RK4.
//
// The integrators are written against a small "system" concept so that the same controller code
// serves a thread-per-trajectory layout (whole state in one thread's registers) and the
// warp-per-trajectory layout of the packed-bed reactor (one axial node per lane): NL local elements,
// rhs(), reduce() (sum over the lanes of the trajectory), global_index(), active().
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include "fdiv.cuh"
#include "jump_models.cuh"   // status bits
#include "ode_models.cuh"

namespace sodm {

constexpr int ODE_MAX_PARAM = 64;

struct OdeLaunch {
    int64_t n_traj, ld_in, ld_out;
    const double* y0;
    int32_t y0_bcast;
    int32_t nx;                    // packed_pfr_ode
    double par[ODE_MAX_PARAM];
    int8_t var_row[ODE_MAX_PARAM];
    const double* var_values;
    double t0;
    int32_t n_out, out_stride, out_last_only, n_kept;
    double rtol, atol, tau;
    void* out;
    int32_t* nfev;
    int32_t* status;
    int64_t max_steps;
    int32_t n_steps_fixed;         // Euler / RK4: number of steps (host replays the FP64 accumulation of t)
};

// ---- Dormand-Prince coefficients (scipy rk.py RK45; Dormand & Prince 1980, Shampine 1986) ------------
#define DP_A21 (1.0 / 5.0)
#define DP_A31 (3.0 / 40.0)
#define DP_A32 (9.0 / 40.0)
#define DP_A41 (44.0 / 45.0)
#define DP_A42 (-56.0 / 15.0)
#define DP_A43 (32.0 / 9.0)
#define DP_A51 (19372.0 / 6561.0)
#define DP_A52 (-25360.0 / 2187.0)
#define DP_A53 (64448.0 / 6561.0)
#define DP_A54 (-212.0 / 729.0)
#define DP_A61 (9017.0 / 3168.0)
#define DP_A62 (-355.0 / 33.0)
#define DP_A63 (46732.0 / 5247.0)
#define DP_A64 (49.0 / 176.0)
#define DP_A65 (-5103.0 / 18656.0)
#define DP_B1 (35.0 / 384.0)
#define DP_B3 (500.0 / 1113.0)
#define DP_B4 (125.0 / 192.0)
#define DP_B5 (-2187.0 / 6784.0)
#define DP_B6 (11.0 / 84.0)
#define DP_E1 (-71.0 / 57600.0)
#define DP_E3 (71.0 / 16695.0)
#define DP_E4 (-71.0 / 1920.0)
#define DP_E5 (17253.0 / 339200.0)
#define DP_E6 (-22.0 / 525.0)
#define DP_E7 (1.0 / 40.0)

// kept grid index of grid point g, or -1
__host__ __device__ __forceinline__ int kept_index(const OdeLaunch& L, int g) {
    if (L.out_last_only) return g == L.n_out - 1 ? 0 : -1;
    if (L.out_stride == 1) return g;                                  // (a runtime integer division is ~25 instructions)
    return (g % L.out_stride) == 0 ? g / L.out_stride : -1;
}

// smallest kept grid index >= g (n_out if there is none): lets the output loops jump over grid points that are not kept
__host__ __device__ __forceinline__ int next_kept(const OdeLaunch& L, int g) {
    if (L.out_last_only) return g <= L.n_out - 1 ? L.n_out - 1 : L.n_out;
    if (L.out_stride == 1) return g < L.n_out ? g : L.n_out;
    const int r = g % L.out_stride;
    const int k = r == 0 ? g : g + (L.out_stride - r);
    return k < L.n_out ? k : L.n_out;
}

template <class Sys>
__host__ __device__ __forceinline__ double rms_scaled(Sys& sys, const double* v, const double* scale) {
    double acc = 0.0;
#pragma unroll
    for (int e = 0; e < Sys::NL; ++e) {
        const double r = sys.active(e) ? div_z(v[e], scale[e]) : 0.0;
        acc += r * r;
    }
    return sqrt(sys.reduce(acc)) / sqrt((double)sys.n_total());
}

template <class Sys, typename OutT>
__host__ __device__ __forceinline__ void store_grid(Sys& sys, const OdeLaunch& L, int64_t i, int kj, const double* v) {
    OutT* out = (OutT*)L.out + i;
#pragma unroll
    for (int e = 0; e < Sys::NL; ++e)
        if (sys.active(e)) out[((int64_t)sys.global_index(e) * L.n_kept + kj) * L.ld_out] = (OutT)v[e];
}

template <class Sys, typename OutT>
__host__ __device__ __forceinline__ void fill_nan_from(Sys& sys, const OdeLaunch& L, int64_t i, int g) {
    double v[Sys::NL];
#pragma unroll
    for (int e = 0; e < Sys::NL; ++e) v[e] = NAN;
    for (; g < L.n_out; ++g) {
        const int kj = kept_index(L, g);
        if (kj >= 0) store_grid<Sys, OutT>(sys, L, i, kj, v);
    }
}

// ------------------------------------------------------------------------------------------------
// scipy-faithful RK45
// ------------------------------------------------------------------------------------------------
// Where the stage derivatives K2 .. K6 of a step live.  K1 (= K7 of the previous step, FSAL) and K7 stay in registers;
// K2 .. K6 are written once per attempt and read back by the later stage combinations, the 5th-order solution, the error
// estimate and -- rarely -- the dense output, so a system with many local elements can park them in shared memory
// (the reactor kernel: 40 of 72 live doubles per lane leave the register file, sodm_ode.cu).  Same arithmetic either way.
template <int NL, int NSLOT = 5>
struct RegisterStages {
    double k[NSLOT][NL];
    __host__ __device__ __forceinline__ double get(int j, int e) const { return k[j][e]; }
    __host__ __device__ __forceinline__ void put(int j, const double* v) {
#pragma unroll
        for (int e = 0; e < NL; ++e) k[j][e] = v[e];
    }
};

// The two explicit pairs of scipy.integrate (`_ivp/rk.py`): what differs between RK45 (Dormand-Prince 5(4), 6 stages) and
// RK23 (Bogacki-Shampine 3(2), 3 stages) is the tableau, the error weights, the dense-output polynomial and the exponent
// 1 / (error_estimator_order + 1) of the step-size controller; everything else -- `select_initial_step`, `_step_impl`, the
// evaluation of the dense output -- is `RungeKutta`'s and is written once below.  K1 = f(t, y) (first same as last) and
// Kl = f(t + h, y_new) live in registers, the stages in between in `S` (slot j = stage j + 2).
struct SchemeRK45 {
    static constexpr int NSLOT = 5;                                // stage vectors kept in `S`
    static constexpr bool LAZY_DENSE = false;                      // the dense output needs nothing beyond the step's own stages
    static constexpr int NQ = 4;                                   // dense output: quartic
    __host__ __device__ static constexpr double exponent() { return 0.2; }      // 1 / (4 + 1)
    template <class Sys, class Stages>
    __host__ __device__ static __forceinline__ void stages(Sys& sys, double t, double h, const double* y, const double* K1, Stages& S,
                                                           double* ynew, double* Kl) {
        constexpr int NL = Sys::NL;
        double kt[NL], ytmp[NL];
#pragma unroll
        for (int e = 0; e < NL; ++e) ytmp[e] = y[e] + (K1[e] * DP_A21) * h;
        sys.rhs(t + (1.0 / 5.0) * h, ytmp, kt); S.put(0, kt);
#pragma unroll
        for (int e = 0; e < NL; ++e) ytmp[e] = y[e] + (K1[e] * DP_A31 + S.get(0, e) * DP_A32) * h;
        sys.rhs(t + (3.0 / 10.0) * h, ytmp, kt); S.put(1, kt);
#pragma unroll
        for (int e = 0; e < NL; ++e) ytmp[e] = y[e] + (K1[e] * DP_A41 + S.get(0, e) * DP_A42 + S.get(1, e) * DP_A43) * h;
        sys.rhs(t + (4.0 / 5.0) * h, ytmp, kt); S.put(2, kt);
#pragma unroll
        for (int e = 0; e < NL; ++e)
            ytmp[e] = y[e] + (K1[e] * DP_A51 + S.get(0, e) * DP_A52 + S.get(1, e) * DP_A53 + S.get(2, e) * DP_A54) * h;
        sys.rhs(t + (8.0 / 9.0) * h, ytmp, kt); S.put(3, kt);
#pragma unroll
        for (int e = 0; e < NL; ++e)
            ytmp[e] = y[e] + (K1[e] * DP_A61 + S.get(0, e) * DP_A62 + S.get(1, e) * DP_A63 + S.get(2, e) * DP_A64 + S.get(3, e) * DP_A65) * h;
        sys.rhs(t + h, ytmp, kt); S.put(4, kt);
#pragma unroll
        for (int e = 0; e < NL; ++e)
            ynew[e] = y[e] + h * (K1[e] * DP_B1 + S.get(1, e) * DP_B3 + S.get(2, e) * DP_B4 + S.get(3, e) * DP_B5 + S.get(4, e) * DP_B6);
        sys.rhs(t + h, ynew, Kl);
    }
    static constexpr int NFEV = 6;
    template <class Stages>
    __host__ __device__ static __forceinline__ double error(const double* K1, const Stages& S, const double* Kl, int e) {
        return K1[e] * DP_E1 + S.get(1, e) * DP_E3 + S.get(2, e) * DP_E4 + S.get(3, e) * DP_E5 + S.get(4, e) * DP_E6 + Kl[e] * DP_E7;
    }
    template <class Stages>
    __host__ __device__ static __forceinline__ void dense_q(const double* K1, const Stages& S, const double* Kl, int e, double* q) {
        q[0] = K1[e];                                             // P[0] = [1, 0, 0, 0] and P[s > 0][0] = 0
        q[1] = K1[e] * (-8048581381.0 / 2820520608.0) + S.get(1, e) * (131558114200.0 / 32700410799.0) +
               S.get(2, e) * (-1754552775.0 / 470086768.0) + S.get(3, e) * (127303824393.0 / 49829197408.0) +
               S.get(4, e) * (-282668133.0 / 205662961.0) + Kl[e] * (40617522.0 / 29380423.0);
        q[2] = K1[e] * (8663915743.0 / 2820520608.0) + S.get(1, e) * (-68118460800.0 / 10900136933.0) +
               S.get(2, e) * (14199869525.0 / 1410260304.0) + S.get(3, e) * (-318862633887.0 / 49829197408.0) +
               S.get(4, e) * (2019193451.0 / 616988883.0) + Kl[e] * (-110615467.0 / 29380423.0);
        q[3] = K1[e] * (-12715105075.0 / 11282082432.0) + S.get(1, e) * (87487479700.0 / 32700410799.0) +
               S.get(2, e) * (-10690763975.0 / 1880347072.0) + S.get(3, e) * (701980252875.0 / 199316789632.0) +
               S.get(4, e) * (-1453857185.0 / 822651844.0) + Kl[e] * (69997945.0 / 29380423.0);
    }
};

// scipy's RK23 (`rk.py`: C = [0, 1/2, 3/4], A = [[1/2], [0, 3/4]], B = [2/9, 1/3, 4/9], E = [5/72, -1/12, -1/9, 1/8],
// P = [[1, -4/3, 5/9], [0, 1, -2/3], [0, 4/3, -8/9], [0, -1, 1]]); zero coefficients are left out as in RK45 (x + 0 K = x).
struct SchemeRK23 {
    static constexpr int NSLOT = 2;
    static constexpr bool LAZY_DENSE = false;
    static constexpr int NQ = 3;                                   // dense output: cubic
    __host__ __device__ static constexpr double exponent() { return 1.0 / 3.0; }      // 1 / (2 + 1)
    template <class Sys, class Stages>
    __host__ __device__ static __forceinline__ void stages(Sys& sys, double t, double h, const double* y, const double* K1, Stages& S,
                                                           double* ynew, double* Kl) {
        constexpr int NL = Sys::NL;
        double kt[NL], ytmp[NL];
#pragma unroll
        for (int e = 0; e < NL; ++e) ytmp[e] = y[e] + (K1[e] * (1.0 / 2.0)) * h;
        sys.rhs(t + (1.0 / 2.0) * h, ytmp, kt); S.put(0, kt);
#pragma unroll
        for (int e = 0; e < NL; ++e) ytmp[e] = y[e] + (S.get(0, e) * (3.0 / 4.0)) * h;
        sys.rhs(t + (3.0 / 4.0) * h, ytmp, kt); S.put(1, kt);
#pragma unroll
        for (int e = 0; e < NL; ++e) ynew[e] = y[e] + h * (K1[e] * (2.0 / 9.0) + S.get(0, e) * (1.0 / 3.0) + S.get(1, e) * (4.0 / 9.0));
        sys.rhs(t + h, ynew, Kl);
    }
    static constexpr int NFEV = 3;
    template <class Stages>
    __host__ __device__ static __forceinline__ double error(const double* K1, const Stages& S, const double* Kl, int e) {
        return K1[e] * (5.0 / 72.0) + S.get(0, e) * (-1.0 / 12.0) + S.get(1, e) * (-1.0 / 9.0) + Kl[e] * (1.0 / 8.0);
    }
    template <class Stages>
    __host__ __device__ static __forceinline__ void dense_q(const double* K1, const Stages& S, const double* Kl, int e, double* q) {
        q[0] = K1[e];
        q[1] = K1[e] * (-4.0 / 3.0) + S.get(0, e) + S.get(1, e) * (4.0 / 3.0) - Kl[e];
        q[2] = K1[e] * (5.0 / 9.0) + S.get(0, e) * (-2.0 / 3.0) + S.get(1, e) * (-8.0 / 9.0) + Kl[e];
    }
};

}  // namespace sodm
#include "dop853_scheme.cuh"      // SchemeDOP853: generated from scipy's coefficient module (tools/gen_dop853.py)
namespace sodm {

// The pieces of scipy's `RungeKutta` (`__init__` -> select_initial_step, `_step_impl`, `RkDenseOutput`), shared by the two
// drivers below: the same arithmetic in the same order whichever driver runs it.
template <class Sys, class Stages, class Scheme = SchemeRK45>
struct Rk45 {
    static constexpr int NL = Sys::NL;
    double K1[NL], K7[NL];           // K1 = f(t, y) (FSAL), K7 = f(t + h, y_new) (the last stage, whatever its number)
    double ynew[NL];
    double rtol, atol, t_bound, t, h_abs, h, t_new;
    int nfev;
    uint32_t st;
    bool dense_ready;                // DOP853: the three extra stages of the last accepted step have been evaluated

    // RungeKutta.__init__: f(t0, y0) and select_initial_step
    __host__ __device__ __forceinline__ void init(Sys& sys, const OdeLaunch& L, const double* y) {
        double kt[NL], ytmp[NL], scale[NL];
        rtol = fmax(L.rtol, 100.0 * DBL_EPSILON);                    // validate_tol
        atol = L.atol;
        t_bound = L.t0 + (double)(L.n_out - 1);
        t = L.t0;
        st = ST_OK;
        nfev = 0;
        h = 0.0;
        t_new = t;
        dense_ready = false;
        sys.rhs(t, y, K1); ++nfev;
        const double interval = fabs(t_bound - t);
        if (interval == 0.0) {
            h_abs = 0.0;
        } else {
#pragma unroll
            for (int e = 0; e < NL; ++e) scale[e] = atol + fabs(y[e]) * rtol;
            const double d0 = rms_scaled(sys, y, scale);
            const double d1 = rms_scaled(sys, K1, scale);
            double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
            h0 = fmin(h0, interval);
#pragma unroll
            for (int e = 0; e < NL; ++e) ytmp[e] = y[e] + h0 * K1[e];
            sys.rhs(t + h0, ytmp, kt); ++nfev;
#pragma unroll
            for (int e = 0; e < NL; ++e) kt[e] = kt[e] - K1[e];
            const double d2 = rms_scaled(sys, kt, scale) / h0;
            const double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / fmax(d1, d2), Scheme::exponent());
            h_abs = fmin(fmin(100.0 * h0, h1), interval);
        }
    }

    // `_step_impl`: attempts until one is accepted.  On success (t, t + h] = (t, t_new] is covered by y -> ynew with the stages in
    // K1, S, K7; returns false when the step size underflows or the error norm is not finite (status bits set).
    __host__ __device__ __forceinline__ bool step(Sys& sys, const double* y, Stages& S) {
        const double min_step = 10.0 * (nextafter(t, (double)INFINITY) - t);   // double overload on host AND device
        if (h_abs < min_step) h_abs = min_step;
        bool rejected = false, failed = false;
        for (;;) {
            if (h_abs < min_step) { failed = true; break; }
            h = h_abs;
            t_new = t + h;
            if (t_new - t_bound > 0.0) t_new = t_bound;
            h = t_new - t;
            h_abs = fabs(h);
            Scheme::stages(sys, t, h, y, K1, S, ynew, K7);         // rk_step
            nfev += Scheme::NFEV;
            // ---- error norm --------------------------------------------------------------------
            double error_norm;
            if constexpr (Scheme::LAZY_DENSE) {                    // DOP853: two estimators (rk.py, _estimate_error_norm)
                error_norm = Scheme::error_norm(sys, y, ynew, K1, S, K7, h, rtol, atol);
            } else {
                double acc = 0.0;
#pragma unroll
                for (int e = 0; e < NL; ++e) {
                    const double sc = atol + fmax(fabs(y[e]), fabs(ynew[e])) * rtol;
                    const double err = Scheme::error(K1, S, K7, e) * h;
                    const double r = sys.active(e) ? div_z(err, sc) : 0.0;        // zero error estimates stay on the fast path (fdiv.cuh)
                    acc += r * r;
                }
                error_norm = sqrt(sys.reduce(acc)) / sqrt((double)sys.n_total());
            }
            if (error_norm < 1.0) {
                double factor = error_norm == 0.0 ? 10.0 : fmin(10.0, 0.9 * pow(error_norm, -Scheme::exponent()));
                if (rejected) factor = fmin(1.0, factor);
                h_abs *= factor;
                dense_ready = false;
                break;
            }
            if (!(error_norm == error_norm)) { st |= ST_NONFINITE; failed = true; break; }
            h_abs *= fmax(0.2, 0.9 * pow(error_norm, -Scheme::exponent()));
            rejected = true;
        }
        if (failed) st |= ST_STEP_TOO_SMALL;
        return !failed;
    }

    // Dense output of the last accepted step at grid time tg in [t, t_new].  RK45 / RK23 (RkDenseOutput):
    // y + h (q0 x + q1 x^2 + ...), x = (tg - t) / h.  DOP853: three more stages on the first call of a step (solve_ivp asks
    // for `dense_output()` only on steps that contain a grid point, and nfev counts them), then Dop853DenseOutput.
    __host__ __device__ __forceinline__ void dense(Sys& sys, const double* y, Stages& S, double tg, double* out) {
        const double x = (tg - t) / h;
        if constexpr (Scheme::LAZY_DENSE) {
            if (!dense_ready) {
                Scheme::prepare_dense(sys, t, h, y, K1, S, K7);
                nfev += 3;
                dense_ready = true;
            }
#pragma unroll
            for (int e = 0; e < NL; ++e) out[e] = Scheme::dense_eval(x, h, y, ynew, K1, S, K7, e);
        } else {
            const double x2 = x * x, x3 = x2 * x, x4 = x3 * x;
#pragma unroll
            for (int e = 0; e < NL; ++e) {
                double q[4];
                q[3] = 0.0;
                Scheme::dense_q(K1, S, K7, e, q);
                if (Scheme::NQ == 4) out[e] = y[e] + h * (q[0] * x + q[1] * x2 + q[2] * x3 + q[3] * x4);
                else out[e] = y[e] + h * (q[0] * x + q[1] * x2 + q[2] * x3);
            }
        }
    }

    // accept the step: y <- y_new, FSAL
    __host__ __device__ __forceinline__ void commit(double* y) {
#pragma unroll
        for (int e = 0; e < NL; ++e) { y[e] = ynew[e]; K1[e] = K7[e]; }
        t = t_new;
    }
};

}  // namespace sodm
#include "implicit.cuh"           // RadauStepper / BdfStepper: scipy's implicit methods behind the same stepper interface
namespace sodm {

// the stepper of a scheme: the shared explicit Runge-Kutta stepper, or one of the implicit ones
template <class Sys, class Stages, class Scheme> struct StepperFor { using type = Rk45<Sys, Stages, Scheme>; };
template <class Sys, class Stages> struct StepperFor<Sys, Stages, SchemeRadau> { using type = RadauStepper<Sys>; };
template <class Sys, class Stages> struct StepperFor<Sys, Stages, SchemeBDF> { using type = BdfStepper<Sys>; };

// Driver 1: every grid point is stored as soon as its step is accepted.  For systems whose lanes move in lockstep (the
// reactor kernel: all lanes of a draw belong to ONE trajectory) and for the CPU harness.
template <class Sys, typename OutT, class Stages, class Scheme = SchemeRK45>
__host__ __device__ __forceinline__ void rk45_run(Sys& sys, const OdeLaunch& L, int64_t i, double* y, Stages& S) {
    constexpr int NL = Sys::NL;
    typename StepperFor<Sys, Stages, Scheme>::type R;
    double ytmp[NL];
    int g = 0;                                                        // next grid point to emit
    R.init(sys, L, y);
    if (R.t == R.t_bound) {                                          // zero-length run: only y0 is output
        const int kj = kept_index(L, 0);
        if (kj >= 0) store_grid<Sys, OutT>(sys, L, i, kj, y);
        g = 1;
    }
    int64_t steps = 0;
    while (R.t < R.t_bound) {
        if (steps++ >= L.max_steps) { R.st |= ST_MAX_STEPS; break; }
        if (!R.step(sys, y, S)) break;
        // ---- dense output for the grid points in (t_old, t_new] (and t0 itself on the first step) ------
        g = next_kept(L, g);
        while (g < L.n_out && L.t0 + (double)g <= R.t_new) {
            R.dense(sys, y, S, L.t0 + (double)g, ytmp);
            store_grid<Sys, OutT>(sys, L, i, kept_index(L, g), ytmp);
            g = next_kept(L, g + 1);
        }
        R.commit(y);
    }
    if (g < L.n_out) fill_nan_from<Sys, OutT>(sys, L, i, g);
    if (sys.leader()) {
        if (L.nfev) L.nfev[i] = R.nfev;
        if (L.status) L.status[i] = (int32_t)R.st;
    }
}

// "every lane of the warp": the device votes; the CPU harness runs the same driver as a warp of one lane
#ifdef __CUDA_ARCH__
__device__ __forceinline__ bool window_all(bool x) { return __all_sync(0xffffffffu, x); }
#else
inline bool window_all(bool x) { return x; }
#endif

// Driver 2 (one trajectory per THREAD): the lanes of a warp run their own step sequences, so with driver 1 every lane
// reaches grid point j at a different moment and its 8-byte store lands alone in a 32-byte sector: ncu on C1 (tutorial SIR,
// 10^6 draws) showed 11.4 GB of DRAM traffic for 2.9 GB of output -- partial-sector writes that L2 completes by reading the
// line back (profiles/r2_sir_ode_ncu_summary.txt).  Here the kept grid points are delivered in WINDOWS of W consecutive
// points: a lane parks its values in shared memory ([point][element][thread]), stops stepping once its window is full, and
// when the whole warp is there every (element, point) row is stored by all lanes at once -- 256 contiguous bytes.  A lane
// that runs ahead waits at most one window; the warp takes max-over-lanes steps either way.  The arithmetic is driver 1's.
template <class Sys, typename OutT, class Stages, class Scheme = SchemeRK45>
__host__ __device__ __forceinline__ void rk45_run_windowed(Sys& sys, const OdeLaunch& L, int64_t i, bool valid, double* y, Stages& S,
                                                  double* win, int W, int stride) {
    constexpr int NL = Sys::NL;
    typename StepperFor<Sys, Stages, Scheme>::type R;
    double ytmp[NL];
    bool pending = false, dead = !valid;
    // the next kept grid point: grid index g, kept index kj (kept points are g = 0, s, 2 s, .. or the last one alone); carried
    // along instead of being recomputed by division
    const int g_step = L.out_last_only ? L.n_out : L.out_stride;
    int g = L.out_last_only ? L.n_out - 1 : 0, kj = 0;
    if (valid) {
        R.init(sys, L, y);
        if (R.t == R.t_bound) {                                      // zero-length run: y0 alone
#pragma unroll
            for (int e = 0; e < NL; ++e) win[e * stride] = y[e];
            kj = L.n_kept;
        }
    } else {
        R.st = ST_OK; R.nfev = 0; R.t = R.t_bound = R.t_new = 0.0; R.h = 1.0;
        kj = L.n_kept;
    }
    int64_t steps = 0;
    int lo = 0;                                                       // first kept index of the current window
    // ONE call site for R.step(): the windows and, behind the last one, the remaining integration of a strided grid that ends
    // before t_bound share the loop (the implicit steppers' step() is thousands of instructions; inlined twice it no longer fits
    // the instruction cache).  `lo` is the same in every lane of the warp, so the vote below is taken by all lanes or by none.
    for (;;) {
        const bool windows = lo < L.n_kept;
        if (!windows && !(valid && !dead && R.t < R.t_bound)) break;  // a strided grid may end before t_bound: the reference's
        bool full = false;                                            // solver still integrates to the end (nfev and status say so)
        if (windows) {
            const int hi = lo + W < L.n_kept ? lo + W : L.n_kept;
            // (a) what the pending step covers of this window
            if (pending) {
                while (kj < hi && L.t0 + (double)g <= R.t_new) {
                    R.dense(sys, y, S, L.t0 + (double)g, ytmp);
                    const int kk = kj - lo;
#pragma unroll
                    for (int e = 0; e < NL; ++e) win[(kk * NL + e) * stride] = ytmp[e];
                    g += g_step; ++kj;
                }
                if (!(kj < L.n_kept && L.t0 + (double)g <= R.t_new)) {   // the step has delivered all its points
                    R.commit(y);
                    pending = false;
                }
            }
            // (b) a trajectory that was abandoned reports NaN from there on (as scipy's failed solve leaves no values)
            if (dead && valid) {
                for (; kj < hi; ++kj, g += g_step) {
                    const int kk = kj - lo;
#pragma unroll
                    for (int e = 0; e < NL; ++e) win[(kk * NL + e) * stride] = NAN;
                }
            }
            // (c) window complete on every lane: store it, one (element, point) row per instruction
            full = kj >= hi;
            if (window_all(full)) {
                if (valid) {
                    OutT* out = (OutT*)L.out + i;
                    for (int kk = 0; kk < hi - lo; ++kk) {
#pragma unroll
                        for (int e = 0; e < NL; ++e)
                            out[((int64_t)sys.global_index(e) * L.n_kept + lo + kk) * L.ld_out] = (OutT)win[(kk * NL + e) * stride];
                    }
                }
                lo = hi;
                continue;
            }
        }
        // (d) the next step
        if (!full && !pending && !dead) {
            if (!(R.t < R.t_bound)) dead = true;                      // (cannot happen: the last step ends on t_bound)
            else if (steps++ >= L.max_steps) { R.st |= ST_MAX_STEPS; dead = true; }
            else if (R.step(sys, y, S)) {
                if (windows) pending = true;
                else R.commit(y);                                     // behind the last window nothing is delivered
            } else dead = true;
        }
    }
    if (valid) {
        if (L.nfev) L.nfev[i] = R.nfev;
        if (L.status) L.status[i] = (int32_t)R.st;
    }
}

// the stages in registers (thread-per-trajectory systems)
template <class Sys, typename OutT, class Scheme = SchemeRK45>
__host__ __device__ __forceinline__ void rk45_run(Sys& sys, const OdeLaunch& L, int64_t i, double* y) {
    RegisterStages<Sys::NL, Scheme::NSLOT> S;
    rk45_run<Sys, OutT, RegisterStages<Sys::NL, Scheme::NSLOT>, Scheme>(sys, L, i, y, S);
}

// ------------------------------------------------------------------------------------------------
// fixed-step schemes + np.interp onto the grid (base.py:800-819)
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ double mul_rn(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}
__host__ __device__ __forceinline__ double add_rn(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}

template <class Sys, typename OutT, bool RK4>
__host__ __device__ __forceinline__ void fixed_step_run(Sys& sys, const OdeLaunch& L, int64_t i, double* y) {
    constexpr int NL = Sys::NL;
    double y_old[NL], k1[NL], k2[NL], k3[NL], k4[NL], ytmp[NL];
    const double tau = L.tau;
    double t = L.t0;
    int g = 1, nfev = 0;
    uint32_t st = ST_OK;
    {
        const int kj = kept_index(L, 0);
        if (kj >= 0) store_grid<Sys, OutT>(sys, L, i, kj, y);
    }
    for (int k = 0; k < L.n_steps_fixed; ++k) {
#pragma unroll
        for (int e = 0; e < NL; ++e) y_old[e] = y[e];
        sys.rhs(t, y, k1); ++nfev;
        if (!RK4) {
#pragma unroll
            for (int e = 0; e < NL; ++e) y[e] = add_rn(y[e], mul_rn(k1[e], tau));     // y_prev + fun(t, y)*tau
        } else {
#pragma unroll
            for (int e = 0; e < NL; ++e) ytmp[e] = y[e] + 0.5 * tau * k1[e];
            sys.rhs(t + 0.5 * tau, ytmp, k2);
#pragma unroll
            for (int e = 0; e < NL; ++e) ytmp[e] = y[e] + 0.5 * tau * k2[e];
            sys.rhs(t + 0.5 * tau, ytmp, k3);
#pragma unroll
            for (int e = 0; e < NL; ++e) ytmp[e] = y[e] + tau * k3[e];
            sys.rhs(t + tau, ytmp, k4);
            nfev += 3;
#pragma unroll
            for (int e = 0; e < NL; ++e) y[e] = y[e] + (tau / 6.0) * (k1[e] + 2.0 * k2[e] + 2.0 * k3[e] + k4[e]);
        }
        const double t_new = t + tau;
        while (g < L.n_out && L.t0 + (double)g <= t_new) {
            const int kj = kept_index(L, g);
            if (kj >= 0) {
                const double tg = L.t0 + (double)g;
                const double t_left = pin_in_loop(t);             // keeps the NL divisions below inside this loop (fdiv.cuh)
#pragma unroll
                for (int e = 0; e < NL; ++e) {
                    if (tg == t_new) ytmp[e] = y[e];
                    else {   // np.interp: slope*(x - xp[j]) + fp[j]
                        const double slope = div_z(y[e] - y_old[e], t_new - t_left);
                        ytmp[e] = add_rn(mul_rn(slope, tg - t_left), y_old[e]);
                    }
                }
                store_grid<Sys, OutT>(sys, L, i, kj, ytmp);
            }
            ++g;
        }
        t = t_new;
    }
    if (sys.leader()) {
        if (L.nfev) L.nfev[i] = nfev;
        if (L.status) L.status[i] = (int32_t)st;
    }
}

// ------------------------------------------------------------------------------------------------
// systems
// ------------------------------------------------------------------------------------------------
template <class M>
struct ThreadSystem {                       // whole state in one thread
    static constexpr int NL = M::N;
    const double* par;
    __host__ __device__ __forceinline__ int n_total() const { return M::N; }
    __host__ __device__ __forceinline__ bool active(int) const { return true; }
    __host__ __device__ __forceinline__ bool leader() const { return true; }
    __host__ __device__ __forceinline__ int global_index(int e) const { return e; }
    __host__ __device__ __forceinline__ double reduce(double x) const { return x; }
    __host__ __device__ __forceinline__ void rhs(double t, const double* y, double* dy) const { M::rhs(t, y, par, dy); }
};

// packed-bed reactor with the whole 8*NX state in one thread (CPU harness only: reference for the
// warp-per-trajectory device layout in sodm_ode.cu)
template <int NX>
struct PfrThreadSystem {
    static constexpr int NL = 8 * NX;
    PfrNode::Coef coef;
    __host__ __device__ __forceinline__ int n_total() const { return NL; }
    __host__ __device__ __forceinline__ bool active(int) const { return true; }
    __host__ __device__ __forceinline__ bool leader() const { return true; }
    __host__ __device__ __forceinline__ int global_index(int e) const { return e; }
    __host__ __device__ __forceinline__ double reduce(double x) const { return x; }
    __host__ __device__ void rhs(double, const double* y, double* dy) const {
        for (int j = 0; j < NX; ++j) {
            double cl[4], cm[4], cr[4], cs[4], dcf[4], dcs[4];
            for (int s = 0; s < 4; ++s) {
                cm[s] = y[s * NX + j];
                cl[s] = j > 0 ? y[s * NX + j - 1] : 0.0;
                cr[s] = j < NX - 1 ? y[s * NX + j + 1] : 0.0;
                cs[s] = y[4 * NX + s * NX + j];
            }
            PfrNode::rhs(coef, j == 0, j == NX - 1, cl, cm, cr, cs, dcf, dcs);
            for (int s = 0; s < 4; ++s) { dy[s * NX + j] = dcf[s]; dy[4 * NX + s * NX + j] = dcs[s]; }
        }
    }
};

template <int NPAR_>
__host__ __device__ __forceinline__ void load_ode_params(const OdeLaunch& L, int64_t i, double* par) {
#pragma unroll
    for (int e = 0; e < NPAR_; ++e) {
        par[e] = L.par[e];
        if (L.var_row[e] >= 0) par[e] = L.var_values[(int64_t)L.var_row[e] * L.ld_in + i];
    }
}

// thread-per-trajectory RK45 with windowed, coalesced output (rk45_run_windowed); `valid` = the thread owns a trajectory
template <class M, typename OutT, class Scheme = SchemeRK45>
__host__ __device__ __forceinline__ void ode_thread_trajectory_windowed(const OdeLaunch& L, int64_t i, bool valid, double* win, int W, int stride) {
    double par[M::NPAR], y[M::N];
    const int64_t src = valid ? i : 0;
    load_ode_params<M::NPAR>(L, src, par);
#pragma unroll
    for (int s = 0; s < M::N; ++s) y[s] = L.y0_bcast ? L.y0[s] : L.y0[(int64_t)s * L.ld_in + src];
    ThreadSystem<M> sys{par};
    RegisterStages<M::N, Scheme::NSLOT> S;
    rk45_run_windowed<ThreadSystem<M>, OutT, RegisterStages<M::N, Scheme::NSLOT>, Scheme>(sys, L, i, valid, y, S, win, W, stride);
}

// one adaptive scheme through driver 1 (direct stores): the implicit methods of systems too large for output windows
template <class M, typename OutT, class Scheme>
__host__ __device__ __forceinline__ void ode_thread_trajectory_scheme(const OdeLaunch& L, int64_t i) {
    double par[M::NPAR], y[M::N];
    load_ode_params<M::NPAR>(L, i, par);
#pragma unroll
    for (int s = 0; s < M::N; ++s) y[s] = L.y0_bcast ? L.y0[s] : L.y0[(int64_t)s * L.ld_in + i];
    ThreadSystem<M> sys{par};
    rk45_run<ThreadSystem<M>, OutT, Scheme>(sys, L, i, y);
}

// WITH_RK45 = false: the kernel of a model whose RK45 runs through the windowed driver carries the fixed-step schemes only
template <class M, typename OutT, bool WITH_RK45 = true>
__host__ __device__ __forceinline__ void ode_thread_trajectory(const OdeLaunch& L, int method, int64_t i) {
    double par[M::NPAR], y[M::N];
    load_ode_params<M::NPAR>(L, i, par);
#pragma unroll
    for (int s = 0; s < M::N; ++s) y[s] = L.y0_bcast ? L.y0[s] : L.y0[(int64_t)s * L.ld_in + i];
    ThreadSystem<M> sys{par};
    if (WITH_RK45 && method == SODM_ODE_RK45) rk45_run<ThreadSystem<M>, OutT>(sys, L, i, y);
    else if (WITH_RK45 && method == SODM_ODE_RK23) rk45_run<ThreadSystem<M>, OutT, SchemeRK23>(sys, L, i, y);
    else if (WITH_RK45 && method == SODM_ODE_DOP853) rk45_run<ThreadSystem<M>, OutT, SchemeDOP853>(sys, L, i, y);
    else if (method == SODM_ODE_EULER) fixed_step_run<ThreadSystem<M>, OutT, false>(sys, L, i, y);
    else fixed_step_run<ThreadSystem<M>, OutT, true>(sys, L, i, y);
}

}  // namespace sodm
