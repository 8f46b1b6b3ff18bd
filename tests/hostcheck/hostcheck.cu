// TEST-ONLY harness: runs the __host__ __device__ trajectory code of pysodm_b200/csrc on the CPU.
//
// The build container has no GPU; this library lets the CPU test-suite check the STATISTICS of the
// very code the kernels execute (binomial sampler, tau-leap / SSA steppers, model functors) against
// the oracle and the reference-generated golden ensembles before any GPU time is spent.  It is not
// part of the product: pysodm_b200 never loads it and has no CPU fallback.
//
// Same argument structs as include/sodm.h, but every "(device)" pointer is a HOST pointer here.
#include <cstring>
#include <vector>
#include "../../pysodm_b200/csrc/jump_setup.cuh"
#include "../../pysodm_b200/csrc/noise.cuh"

using namespace sodm;

template <class M, typename OutT>
static void run_host(const JumpLaunch& L, int method, bool aligned) {
    std::vector<float> table((size_t)L.n_tdpf_values * M::TDPF_COUNT + 1);
    for (size_t e = 0; e + 1 < table.size(); ++e) table[e] = (float)L.tdpf_table[e];
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < L.n_traj; ++i) {
        int wn[M::NDRAW];
        float wp[M::NDRAW];
        uint8_t ord[M::NDRAW];
        uint16_t queue[2 * M::NDRAW];
        WorkList w{wn, wp, ord, queue, 1, 0, 1u};
        if (method == SODM_JUMP_TAU_LEAP) {
            const bool two_pass = (L.compat & TUNE_QUEUE) != 0u;
            if (aligned) { if (two_pass) tau_leap_trajectory<M, OutT, true, true>(L, i, true, table.data(), w);
                           else          tau_leap_trajectory<M, OutT, true, false>(L, i, true, table.data(), w); }
            else         { if (two_pass) tau_leap_trajectory<M, OutT, false, true>(L, i, true, table.data(), w);
                           else          tau_leap_trajectory<M, OutT, false, false>(L, i, true, table.data(), w); }
        } else {
            ssa_trajectory<M, OutT>(L, i);
        }
    }
}

template <class M>
static int run_host_dtype(const JumpLaunch& L, int method, bool aligned, int dtype) {
    switch (dtype) {
        case SODM_I32: run_host<M, int32_t>(L, method, aligned); return SODM_OK;
        case SODM_F32: run_host<M, float>(L, method, aligned); return SODM_OK;
        case SODM_F64: run_host<M, double>(L, method, aligned); return SODM_OK;
    }
    return SODM_EINVAL;
}

extern "C" int hc_jump_run(const sodm_jump_args* a) {
    JumpLaunch L;
    bool aligned = false;
    const int rc = build_jump_launch(a, L, aligned);
    if (rc != SODM_OK) return rc;
    switch (a->model) {
        case MODEL_SIR_JUMP:       return run_host_dtype<SirJump>(L, a->method, aligned, a->out_dtype);
        case MODEL_SIR_SI_JUMP:    return run_host_dtype<SirSiJump>(L, a->method, aligned, a->out_dtype);
        case MODEL_INFLUENZA_JUMP: return run_host_dtype<InfluenzaJump>(L, a->method, aligned, a->out_dtype);
    }
    return SODM_EMODEL;
}

// `count` variates of Bin(n, p) from the stream (seed, traj); attempts[0] receives the number of tries.
extern "C" void hc_binomial(uint64_t seed, uint64_t traj, int32_t n, float p, int64_t count, int32_t* out,
                            int64_t* attempts) {
    PhiloxWords rng;
    rng.init(seed, traj);
    int64_t tries = 0;
    for (int64_t c = 0; c < count; ++c) {
        int k = 0;
        do { ++tries; } while (!binomial_try(rng, n, p, k));
        out[c] = k;
    }
    if (attempts) attempts[0] = tries;
}

// ONE draw at a time through the two-pass sampler's unit of work (binomial_attempt): the first attempt without the exact
// test; an undecided candidate repeats the same words with it (as pass 2 does); a rejected one takes fresh words.
// `force_word` != 0 replaces the FIRST u-word (edge cases: all-ones word -> u01 at its top end).
extern "C" void hc_binomial_words(uint64_t seed, uint64_t traj, int32_t n, float p, int64_t count, int32_t* out,
                                  int64_t* attempts, uint32_t force_word) {
    int64_t tries = 0;
    for (int64_t c = 0; c < count; ++c) {
        AttemptKey key;
        key.set(seed, traj, (uint32_t)c);
        int k = 0, state = BIN_RETRY;
        for (uint32_t a = 0; state != BIN_DONE; ++a) {
            ++tries;
            uint32_t wu, wv;
            key.words(0, 3, a % 64u, wu, wv);
            if (force_word && c == 0 && a == 0) wu = force_word;
            state = binomial_attempt<false>(n, p, wu, wv, k);
            if (state == BIN_EXACT) state = binomial_attempt<true>(n, p, wu, wv, k);
        }
        out[c] = k;
    }
    if (attempts) attempts[0] = tries;
}

extern "C" void hc_philox2x32(uint32_t key, uint32_t c0, uint32_t c1, uint32_t* out) {
    philox2x32_10(key, c0, c1, out[0], out[1]);
}

extern "C" void hc_philox(uint64_t seed, uint64_t traj, uint32_t purpose, int64_t n_blocks, uint32_t* out) {
    Philox4x32 g;
    g.init(seed, traj, purpose);
    for (int64_t b = 0; b < n_blocks; ++b) {
        uint32_t w[4];
        g.next4(w);
        std::memcpy(out + 4 * b, w, sizeof(w));
    }
}

extern "C" double hc_binomial_log_ratio_large(double n, double p, double k) {
    double out = 0.0;
    return binomial_log_ratio_large(n, p, floor((n + 1.0) * p), k, out) ? out : NAN;
}

extern "C" double hc_binomial_log_ratio(double n, double p, double k) {
    return binomial_log_ratio(n, p, floor((n + 1.0) * p), k);
}

// ---------------------------------------------------------------------------------------------------
// ODE integrators on the CPU (same rk45_run / fixed_step_run code as the kernels)
// ---------------------------------------------------------------------------------------------------
#include "../../pysodm_b200/csrc/ode_setup.cuh"

// `window` > 0: the WINDOWED output driver of the device's thread kernel (rk45_run_windowed) as a warp of one lane -- the
// window bookkeeping (strided grids, partial last window, abandoned runs, the tail integration) is then testable without a GPU
template <class M, typename OutT>
static void ode_host(const OdeLaunch& L, int method, int window) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < L.n_traj; ++i) {
        if (window > 0 && (method == SODM_ODE_RK45 || method == SODM_ODE_RK23 || method == SODM_ODE_DOP853)) {
            std::vector<double> win((size_t)window * M::N);
            if (method == SODM_ODE_RK45) ode_thread_trajectory_windowed<M, OutT, SchemeRK45>(L, i, true, win.data(), window, 1);
            else if (method == SODM_ODE_RK23) ode_thread_trajectory_windowed<M, OutT, SchemeRK23>(L, i, true, win.data(), window, 1);
            else ode_thread_trajectory_windowed<M, OutT, SchemeDOP853>(L, i, true, win.data(), window, 1);
        } else if (method == SODM_ODE_RADAU || method == SODM_ODE_BDF) {        // implicit.cuh; window > 0: through the windowed driver
            std::vector<double> win((size_t)(window > 0 ? window : 1) * M::N);
            if (window > 0 && method == SODM_ODE_RADAU) ode_thread_trajectory_windowed<M, OutT, SchemeRadau>(L, i, true, win.data(), window, 1);
            else if (window > 0) ode_thread_trajectory_windowed<M, OutT, SchemeBDF>(L, i, true, win.data(), window, 1);
            else if (method == SODM_ODE_RADAU) ode_thread_trajectory_scheme<M, OutT, SchemeRadau>(L, i);
            else ode_thread_trajectory_scheme<M, OutT, SchemeBDF>(L, i);
        } else {
            ode_thread_trajectory<M, OutT>(L, method, i);
        }
    }
}

template <int NX, typename OutT>
static void pfr_host(const OdeLaunch& L, int method) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < L.n_traj; ++i) {
        double par[PfrNode::NPAR];
        load_ode_params<PfrNode::NPAR>(L, i, par);
        static thread_local double y[8 * NX];
        for (int s = 0; s < 8 * NX; ++s) y[s] = L.y0_bcast ? L.y0[s] : L.y0[(int64_t)s * L.ld_in + i];
        PfrThreadSystem<NX> sys{PfrNode::coefficients(par)};
        if (method == SODM_ODE_RK45) rk45_run<PfrThreadSystem<NX>, OutT>(sys, L, i, y);
        else if (method == SODM_ODE_RK23) rk45_run<PfrThreadSystem<NX>, OutT, SchemeRK23>(sys, L, i, y);
        else if (method == SODM_ODE_DOP853) rk45_run<PfrThreadSystem<NX>, OutT, SchemeDOP853>(sys, L, i, y);
        else if (method == SODM_ODE_EULER) fixed_step_run<PfrThreadSystem<NX>, OutT, false>(sys, L, i, y);
        else fixed_step_run<PfrThreadSystem<NX>, OutT, true>(sys, L, i, y);
    }
}

// Test-only stiff systems for the implicit methods (tests/golden/make_golden.py::gen_ode_stiff defines the same right-hand sides,
// operation for operation, as models of the unmodified reference): a->model 0 = Robertson's kinetics, 1 = Van der Pol.
struct RobertsonOde {
    static constexpr int N = 3, NPAR = 3;
    __host__ __device__ static void rhs(double, const double* y, const double* par, double* dy) {
        const double r1 = par[0] * y[0];
        const double r2 = par[1] * y[1] * y[2];
        const double r3 = par[2] * y[1] * y[1];
        dy[0] = -r1 + r2;
        dy[1] = r1 - r2 - r3;
        dy[2] = r3;
    }
};
struct VanDerPolOde {
    static constexpr int N = 2, NPAR = 1;
    __host__ __device__ static void rhs(double, const double* y, const double* par, double* dy) {
        dy[0] = y[1];
        dy[1] = par[0] * (1 - y[0] * y[0]) * y[1] - y[0];
    }
};

// The dense LU of the implicit steppers (implicit.cuh) on a caller's matrix: n = 3 takes the compile-time-indexed path, n = 12 the
// run-time one; `aim` != nullptr: complex.  Solves A x = b in place (b, bim); A is row-major.
template <int N>
static void lu_host(const double* are, const double* aim, double* bre, double* bim) {
    if (aim) {
        LuComplex<N> lu;
        for (int i = 0; i < N * N; ++i) { lu.re[i] = are[i]; lu.im[i] = aim[i]; }
        lu.factor();
        lu.solve(bre, bim);
    } else {
        LuReal<N> lu;
        for (int i = 0; i < N * N; ++i) lu.a[i] = are[i];
        lu.factor();
        lu.solve(bre);
    }
}
// BdfStepper::change_D on a caller's difference array D[8][3] (rows above `order` must stay untouched)
extern "C" void hc_bdf_change_D(int order, double factor, double* D) {
    BdfStepper<ThreadSystem<SirOde>> R;
    for (int k = 0; k < 8; ++k)
        for (int e = 0; e < 3; ++e) R.D[k][e] = D[k * 3 + e];
    R.change_D(order, factor);
    for (int k = 0; k < 8; ++k)
        for (int e = 0; e < 3; ++e) D[k * 3 + e] = R.D[k][e];
}

extern "C" int hc_lu_solve(int n, const double* are, const double* aim, double* bre, double* bim) {
    if (n == 3) lu_host<3>(are, aim, bre, bim);
    else if (n == 8) lu_host<8>(are, aim, bre, bim);
    else if (n == 12) lu_host<12>(are, aim, bre, bim);
    else return SODM_EUNSUPPORTED;
    return SODM_OK;
}

// a->model 2 / 3: a chain  y_i' = a_i y_(i-1) + b_i y_i + c_i y_(i+1) - q y_i^2  of 3 / 12 equations (par = a[N], b[N], c[N], q):
// random stiff systems for the implicit steppers against scipy itself (tests/test_hostcheck.py evaluates the same expression,
// operation for operation, in Python); 12 equations take the run-time-indexed path of implicit.cuh
template <int N_>
struct ChainOde {
    static constexpr int N = N_, NPAR = 3 * N_ + 1;
    __host__ __device__ static void rhs(double, const double* y, const double* par, double* dy) {
        for (int i = 0; i < N; ++i) {
            double s = par[N + i] * y[i];
            if (i > 0) s = par[i] * y[i - 1] + s;
            if (i < N - 1) s = s + par[2 * N + i] * y[i + 1];
            dy[i] = s - par[3 * N] * y[i] * y[i];
        }
    }
};

// ImplicitBase::num_jac of a test functor at (t, y): `factor` (in/out; `have_factor` = 0 starts from scipy's default) and J (row-major)
template <class M>
static void num_jac_host(const double* par, double t, const double* y, double atol, int have_factor, double* factor, double* J) {
    ThreadSystem<M> sys{par};
    ImplicitBase<ThreadSystem<M>> B;
    B.atol = atol;
    B.have_factor = have_factor != 0;
    for (int j = 0; j < M::N; ++j) B.jac_factor[j] = factor[j];
    double f[M::N];
    sys.rhs(t, y, f);
    B.num_jac(sys, t, y, f);
    for (int j = 0; j < M::N; ++j) factor[j] = B.jac_factor[j];
    for (int i = 0; i < M::N * M::N; ++i) J[i] = B.J[i];
}
extern "C" int hc_num_jac(int model, const double* par, double t, const double* y, double atol, int have_factor, double* factor, double* J) {
    if (model == 0) num_jac_host<RobertsonOde>(par, t, y, atol, have_factor, factor, J);
    else if (model == 1) num_jac_host<VanDerPolOde>(par, t, y, atol, have_factor, factor, J);
    else if (model == 2) num_jac_host<ChainOde<3>>(par, t, y, atol, have_factor, factor, J);
    else if (model == 3) num_jac_host<ChainOde<12>>(par, t, y, atol, have_factor, factor, J);
    else return SODM_EMODEL;
    return SODM_OK;
}

extern "C" int hc_ode_test_model_run(const sodm_ode_args* a) {
    OdeLaunch L;
    sodm_model_info info = {};
    info.kind = SODM_KIND_ODE;
    const int ns[4] = {RobertsonOde::N, VanDerPolOde::N, 3, 12}, np_[4] = {RobertsonOde::NPAR, VanDerPolOde::NPAR, 10, 37};
    if (a->model < 0 || a->model > 3) return SODM_EMODEL;
    info.n_state = ns[a->model];
    info.n_param = np_[a->model];
    const int rc = build_ode_launch_for(a, L, info);
    if (rc != SODM_OK) return rc;
    if (a->model == 0) ode_host<RobertsonOde, double>(L, a->method, a->reserved);
    else if (a->model == 1) ode_host<VanDerPolOde, double>(L, a->method, a->reserved);
    else if (a->model == 2) ode_host<ChainOde<3>, double>(L, a->method, a->reserved);
    else ode_host<ChainOde<12>, double>(L, a->method, a->reserved);
    return SODM_OK;
}

extern "C" int hc_ode_run(const sodm_ode_args* a) {
    OdeLaunch L;
    sodm_model_info info;
    const int rc = build_ode_launch(a, L, info);
    if (rc != SODM_OK) return rc;
    if (a->out_dtype != SODM_F64) return SODM_EUNSUPPORTED;
    switch (a->model) {
        // sodm_ode_args.reserved (unused by the product) carries the window length of the harness's windowed runs
        case MODEL_SIR_ODE:      ode_host<SirOde, double>(L, a->method, a->reserved); return SODM_OK;
        case MODEL_SIR_RATE_ODE: ode_host<SirRateOde, double>(L, a->method, a->reserved); return SODM_OK;
        case MODEL_SIR_SI_ODE:   ode_host<SirSiOde, double>(L, a->method, a->reserved); return SODM_OK;
        case MODEL_PPBB_ODE:     ode_host<PpbbOde, double>(L, a->method, a->reserved); return SODM_OK;
        case MODEL_PACKED_PFR_ODE:
            if (a->method == SODM_ODE_RADAU || a->method == SODM_ODE_BDF) return SODM_EUNSUPPORTED;
            if (L.nx == 25) { pfr_host<25, double>(L, a->method); return SODM_OK; }
            if (L.nx == 100) { pfr_host<100, double>(L, a->method); return SODM_OK; }
            return SODM_EUNSUPPORTED;
    }
    return SODM_EMODEL;
}

// `count` noisy values around model output x, streams (seed, traj = 0..count-1, cell) exactly as the sodm_noise kernels use them.
extern "C" int32_t hc_noise(int32_t kind, double param, double x, uint64_t seed, uint32_t cell, int64_t count, double* out) {
    int32_t status = 0;
    for (int64_t i = 0; i < count; ++i) {
        if (kind == SODM_NOISE_GAUSSIAN_ABS || kind == SODM_NOISE_GAUSSIAN_REL) {
            float z[4];
            normal_quad(seed, (uint64_t)i, cell / 4, z);
            out[i] = gaussian_noise(kind == SODM_NOISE_GAUSSIAN_REL, param, x, z[cell % 4]);
            continue;
        }
        ElementRng rng;
        rng.init(seed, (uint64_t)i, cell);
        out[i] = kind == SODM_NOISE_POISSON ? count_noise<false>(rng, param, x, status) : count_noise<true>(rng, param, x, status);
    }
    return status;
}
