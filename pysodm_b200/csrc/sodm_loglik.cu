// Batched log-likelihood of simulated trajectories against one aligned dataset.
//
// Replaces, for a whole population of trajectories at once, the tail of
// `log_posterior_probability.compute_log_likelihood`
// (/root/reference/src/pySODM/optimization/objective_functions.py:250-288: sum over the model dimensions
// absent from the data, linear interpolation in time to the data's time points, selection of the data's
// coordinates) and the `ll_*` function it calls (:550-666).
//
// Layout decision: trajectories are stored [row][time][trajectory], so ONE THREAD PER TRAJECTORY reads
// coalesced (consecutive lanes = consecutive trajectories) and its sum over the n_t x n_col data points is
// thread-local -- no cross-lane reduction is needed for it.  The terms that depend on the data only
// (lnGamma's, sigma terms) are identical for every trajectory: each block evaluates them ONCE,
// distributed over its lanes and combined with a warp-shuffle + shared-memory reduction, instead of
// paying ~50 FP64 lgamma per trajectory.
#include "common.cuh"

namespace sodm {

constexpr int LL_BLOCK = 128;

struct LoglikLaunch {
    int32_t kind, n_out, n_col, n_t, arg_mode;
    int64_t n_traj, ld;
    const void* traj;
    const int32_t* col_ptr;
    const int32_t* col_rows;
    const double* col_weights;
    double t0;
    const double* t_data;
    const double* ydata;
    const double* ll_arg;
    double weight;
    double* lp;
    int32_t* status;
};

__device__ __forceinline__ double ll_arg_at(const LoglikLaunch& L, int k, int c) {
    switch (L.arg_mode) {
        case SODM_ARG_SCALAR: return L.ll_arg[0];
        case SODM_ARG_COLUMN: return L.ll_arg[c];
        case SODM_ARG_FULL:   return L.ll_arg[(int64_t)k * L.n_col + c];
    }
    return 0.0;
}

// Data-only part of the log-likelihood of point (k, c).
__device__ __forceinline__ double data_term(const LoglikLaunch& L, int k, int c) {
    const double yd = L.ydata[(int64_t)k * L.n_col + c];
    switch (L.kind) {
        case SODM_LL_POISSON:                                   // - lnGamma(yd + 1)   (yd already shifted by +1: :631-633)
            return -lgamma(yd + 1.0);
        case SODM_LL_NEGBIN: {                                  // :663-666 with yd' = yd + 1
            const double alpha = ll_arg_at(L, k, c);
            const double ydp = yd + 1.0;
            // last term of :666 is  sum over alpha's OWN elements of n_t * lnGamma(1/alpha)  (SURVEY Q14): once per
            // column for per-column alpha, ONCE in total for a scalar alpha, n_t times per point for per-point alpha
            double w_last = 1.0;
            if (L.arg_mode == SODM_ARG_SCALAR) w_last = (c == 0) ? 1.0 : 0.0;
            else if (L.arg_mode == SODM_ARG_FULL) w_last = (double)L.n_t;
            return ydp * log(alpha) + lgamma(ydp + 1.0 / alpha) - lgamma(ydp + 1.0) - w_last * lgamma(1.0 / alpha);
        }
        case SODM_LL_NORMAL: {                                  // -1/2 ln(2 pi sigma^2)   (:610)
            const double sigma = ll_arg_at(L, k, c);
            return -0.5 * log(2.0 * M_PI * sigma * sigma);
        }
        case SODM_LL_LOGNORMAL: {                               // -1/2 [ln(2 pi sigma^2) + ln(yd + 1)]   (:579)
            const double sigma = ll_arg_at(L, k, c);
            return -0.5 * (log(2.0 * M_PI * sigma * sigma) + log(yd + 1.0));
        }
    }
    return 0.0;
}

// Model-dependent part of point (k, c) given the aligned model prediction ym.
__device__ __forceinline__ double model_term(const LoglikLaunch& L, int k, int c, double ym) {
    const double yd = L.ydata[(int64_t)k * L.n_col + c];
    switch (L.kind) {
        case SODM_LL_POISSON: {                                 // -(ym+1) + (yd+1) ln(ym+1)
            const double ymp = ym + 1.0;
            return -ymp + (yd + 1.0) * log(ymp);
        }
        case SODM_LL_NEGBIN: {                                  // (yd+1) ln(ym+1) - (yd+1 + 1/alpha) ln(1 + alpha (ym+1))
            const double alpha = ll_arg_at(L, k, c);
            const double ymp = ym + 1.0, ydp = yd + 1.0;
            return ydp * log(ymp) - (ydp + 1.0 / alpha) * log(1.0 + alpha * ymp);
        }
        case SODM_LL_NORMAL: {
            const double sigma = ll_arg_at(L, k, c);
            const double d = yd - ym;
            return -0.5 * d * d / (sigma * sigma);
        }
        case SODM_LL_LOGNORMAL: {
            const double sigma = ll_arg_at(L, k, c);
            const double d = log(yd + 1.0) - log(ym + 1.0);
            return -0.5 * d * d / (sigma * sigma);
        }
    }
    return 0.0;
}

template <typename T>
__global__ void __launch_bounds__(LL_BLOCK) loglik_kernel(const __grid_constant__ LoglikLaunch L) {
    __shared__ double warp_part[LL_BLOCK / 32];
    __shared__ double data_const;
    // ---- data-only terms: once per block, spread over the lanes, shuffle-reduced ---------------------
    const int n_pts = L.n_t * L.n_col;
    double part = 0.0;
    for (int q = threadIdx.x; q < n_pts; q += LL_BLOCK) part += data_term(L, q / L.n_col, q % L.n_col);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < LL_BLOCK / 32; ++w) s += warp_part[w];
        data_const = s;
    }
    __syncthreads();

    const int64_t i = (int64_t)blockIdx.x * LL_BLOCK + threadIdx.x;
    if (i >= L.n_traj) return;
    const T* traj = (const T*)L.traj + i;
    double acc = 0.0;
    bool bad = false;
    for (int k = 0; k < L.n_t; ++k) {
        // linear interpolation on the unit-spaced output grid (xarray .interp(method="linear"), :254)
        const double x = L.t_data[k] - L.t0;
        const int j = (int)floor(x);
        const double frac = x - (double)j;
        const bool inside = x >= 0.0 && j + (frac > 0.0 ? 1 : 0) <= L.n_out - 1;
        for (int c = 0; c < L.n_col; ++c) {
            double ym = 0.0;
            if (!inside) ym = NAN;                                     // xarray fills NaN outside the simulated range
            else {
                for (int q = L.col_ptr[c]; q < L.col_ptr[c + 1]; ++q) {
                    const int64_t row = (int64_t)L.col_rows[q] * L.n_out;
                    const double lo = (double)traj[(row + j) * L.ld];
                    double v = lo;
                    if (frac > 0.0) {
                        const double hi = (double)traj[(row + j + 1) * L.ld];
                        v = lo + (hi - lo) * frac;
                    }
                    ym += L.col_weights ? L.col_weights[q] * v : v;
                }
            }
            if (!(ym == ym)) bad = true;                               // :264-265 / :282-283 raise on NaN
            acc += model_term(L, k, c, ym);
        }
    }
    L.lp[i] += L.weight * (acc + data_const);
    if (bad && L.status) L.status[i] |= SODM_ST_NONFINITE;
}

}  // namespace sodm

using namespace sodm;

extern "C" int sodm_loglik(const sodm_loglik_args* a) {
    if (!a) return fail(SODM_EINVAL, "sodm_loglik: null args");
    if (a->kind < SODM_LL_POISSON || a->kind > SODM_LL_NEGBIN) return fail(SODM_EINVAL, "sodm_loglik: unknown kind %d", a->kind);
    if (a->n_traj < 0 || a->ld < a->n_traj) return fail(SODM_EINVAL, "sodm_loglik: need 0 <= n_traj <= ld");
    if (a->n_out < 1 || a->n_col < 1 || a->n_t < 1) return fail(SODM_EINVAL, "sodm_loglik: n_out, n_col, n_t must be >= 1");
    if (a->n_traj == 0) return SODM_OK;
    if (!a->traj || !a->col_ptr || !a->col_rows || !a->t_data || !a->ydata || !a->lp)
        return fail(SODM_EINVAL, "sodm_loglik: traj, col_ptr, col_rows, t_data, ydata and lp are required");
    const bool needs_arg = a->kind != SODM_LL_POISSON;
    if (needs_arg && (a->arg_mode < SODM_ARG_SCALAR || a->arg_mode > SODM_ARG_FULL || !a->ll_arg))
        return fail(SODM_EINVAL, "sodm_loglik: this likelihood needs ll_arg (sigma / alpha) and a valid arg_mode");
    if (a->n_traj == 0) return SODM_OK;
    LoglikLaunch L;
    L.kind = a->kind; L.n_out = a->n_out; L.n_col = a->n_col; L.n_t = a->n_t;
    L.arg_mode = needs_arg ? a->arg_mode : SODM_ARG_NONE;
    L.n_traj = a->n_traj; L.ld = a->ld; L.traj = a->traj;
    L.col_ptr = a->col_ptr; L.col_rows = a->col_rows; L.col_weights = a->col_weights;
    L.t0 = a->t0; L.t_data = a->t_data; L.ydata = a->ydata; L.ll_arg = a->ll_arg;
    L.weight = a->weight; L.lp = a->lp; L.status = a->status;
    const unsigned grid = (unsigned)((a->n_traj + LL_BLOCK - 1) / LL_BLOCK);
    cudaStream_t stream = (cudaStream_t)a->stream;
    switch (a->traj_dtype) {
        case SODM_I32: loglik_kernel<int32_t><<<grid, LL_BLOCK, 0, stream>>>(L); break;
        case SODM_F32: loglik_kernel<float><<<grid, LL_BLOCK, 0, stream>>>(L); break;
        case SODM_F64: loglik_kernel<double><<<grid, LL_BLOCK, 0, stream>>>(L); break;
        default: return fail(SODM_EINVAL, "sodm_loglik: unknown traj_dtype %d", a->traj_dtype);
    }
    SODM_CUDA_CHECK(cudaGetLastError());
    return SODM_OK;
}
