// Optimiser populations on the device: what the reference's optimisers do to 10^2 .. 10^5 parameter
// vectors on the host between two evaluations of the objective, done where the objective runs, so that an iteration is
// "kernels + a handful of numbers back" instead of NumPy passes over the population that cost as much as the simulation:
//   swarm update        v = w v + phi_p r_p (p - x) + phi_g r_g (g - x);  x += v;  clamp to the bounds
//                                           /root/reference/src/pySODM/optimization/pso.py:207-216
//   bests               p, fp <- x, fx where fx < fp;  arg min fp      pso.py:229-236
//   KDE log-density     the proposal density of emcee's KDEMove, which the reference's default cocktail uses with weight 1/4
//                                           /root/reference/src/pySODM/optimization/mcmc.py:26
// Layout: structure-of-arrays [n_dim][ld] FP64, member i of dimension d at base[d * ld + i] -- the layout the simulators
// read their per-trajectory parameters in (`var_values`), so a population feeds `sodm_jump_run` / `sodm_ode_run` as is.
// Random numbers: Philox keyed by (seed, global member id, iteration): independent of launch geometry and sharding.
#include "common.cuh"
#include "philox.cuh"

namespace sodm {

__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {        // uniform on [0, 1) with 53 random bits
    return (double)((((uint64_t)hi << 32) | lo) >> 11) * 1.1102230246251565e-16;
}

struct Bounds {
    double lb[SODM_POPULATION_MAX_DIM], ub[SODM_POPULATION_MAX_DIM];
};

// One thread per particle; every access is a coalesced column read / write.  HBM-bound: 5 reads + 2 writes per element.
__global__ void __launch_bounds__(256) pso_move_kernel(int n_dim, int64_t n, int64_t ld, double* __restrict__ x,
                                                       double* __restrict__ v, const double* __restrict__ p,
                                                       const double* __restrict__ g, Bounds b, double omega, double phip,
                                                       double phig, uint64_t seed, uint32_t iteration, int64_t id0) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        Philox4x32 rng;
        rng.init(seed, (uint64_t)(id0 + i), iteration);
        for (int d = 0; d < n_dim; ++d) {
            uint32_t w[4];
            rng.next4(w);
            const double rp = u53(w[0], w[1]), rg = u53(w[2], w[3]);
            const int64_t k = d * ld + i;
            const double xd = x[k];
            const double vd = omega * v[k] + phip * rp * (p[k] - xd) + phig * rg * (g[d] - xd);
            double xn = xd + vd;
            xn = xn < b.lb[d] ? b.lb[d] : (xn > b.ub[d] ? b.ub[d] : xn);          // pso.py:213-216 (NaN stays NaN)
            v[k] = vd;
            x[k] = xn;
        }
    }
}

__global__ void __launch_bounds__(256) pso_keep_kernel(int n_dim, int64_t n, int64_t ld, const double* __restrict__ x,
                                                       const double* __restrict__ fx, const uint8_t* __restrict__ feasible,
                                                       double* __restrict__ p, double* __restrict__ fp) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double f = fx[i];
        if (f < fp[i] && (!feasible || feasible[i])) {                             // pso.py:229-231
            fp[i] = f;
            for (int d = 0; d < n_dim; ++d) p[d * ld + i] = x[d * ld + i];
        }
    }
}

// np.argmin(fp): the FIRST index of the smallest value, NaN counting as smallest (NumPy's rule).  One CTA: the
// population is at most a few MB, a single pass at L2 speed.
__global__ void __launch_bounds__(1024) pso_argmin_kernel(int n_dim, int64_t n, int64_t ld, const double* __restrict__ p,
                                                          const double* __restrict__ fp, double* __restrict__ best,
                                                          int64_t* __restrict__ best_index) {
    __shared__ double s_val[32];
    __shared__ int64_t s_idx[32];
    double val = 0.0;
    int64_t idx = -1;
    auto better = [](double a, int64_t ia, double b, int64_t ib) {                // is (a, ia) ahead of (b, ib)?
        if (ib < 0) return ia >= 0;
        if (ia < 0) return false;
        const bool an = a != a, bn = b != b;
        if (an || bn) return an && (!bn || ia < ib);
        return a < b || (a == b && ia < ib);
    };
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const double f = fp[i];
        if (better(f, i, val, idx)) { val = f; idx = i; }
    }
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, val, off);
        const int64_t oi = __shfl_down_sync(0xffffffffu, idx, off);
        if (better(ov, oi, val, idx)) { val = ov; idx = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = val; s_idx[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int n_warps = blockDim.x >> 5;
        val = threadIdx.x < n_warps ? s_val[threadIdx.x] : 0.0;
        idx = threadIdx.x < n_warps ? s_idx[threadIdx.x] : -1;
        for (int off = 16; off > 0; off >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, val, off);
            const int64_t oi = __shfl_down_sync(0xffffffffu, idx, off);
            if (better(ov, oi, val, idx)) { val = ov; idx = oi; }
        }
        if (threadIdx.x == 0) { s_val[0] = val; s_idx[0] = idx; }
    }
    __syncthreads();
    idx = s_idx[0];
    if (threadIdx.x == 0) { best[0] = s_val[0]; *best_index = idx; }
    if ((int)threadIdx.x < n_dim) best[1 + threadIdx.x] = p[threadIdx.x * ld + idx];
}

// log sum_k exp(-|q - x_k|^2 / 2) for whitened points: one thread per query, the kernels staged through shared memory
// in tiles every thread of the CTA sweeps (broadcast reads).  Two sweeps: the smallest squared distance first, so that
// the second costs ONE exponential per pair and cannot underflow.  FP64-ALU bound: D subtract + D FMA + exp per pair.
template <int D>
__global__ void __launch_bounds__(128) kde_logpdf_kernel(const double* __restrict__ points, int64_t n_points, int64_t ld_p,
                                                         const double* __restrict__ queries, int64_t n_queries, int64_t ld_q,
                                                         double log_norm, double* __restrict__ out) {
    constexpr int TILE = 256;
    __shared__ double tile[D][TILE];
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double q[D];
#pragma unroll
    for (int d = 0; d < D; ++d) q[d] = j < n_queries ? queries[d * ld_q + j] : 0.0;
    double d_min = 1.0 / 0.0, sum = 0.0;
    for (int pass = 0; pass < 2; ++pass) {
        for (int64_t k0 = 0; k0 < n_points; k0 += TILE) {
            const int m = (int)((n_points - k0) < TILE ? (n_points - k0) : TILE);
            __syncthreads();
            for (int t = threadIdx.x; t < m; t += blockDim.x) {
#pragma unroll
                for (int d = 0; d < D; ++d) tile[d][t] = points[d * ld_p + k0 + t];
            }
            __syncthreads();
            if (pass == 0) {
                for (int t = 0; t < m; ++t) {
                    double r2 = 0.0;
#pragma unroll
                    for (int d = 0; d < D; ++d) { const double e = q[d] - tile[d][t]; r2 = fma(e, e, r2); }
                    d_min = r2 < d_min ? r2 : d_min;                               // NaN distances never win
                }
            } else {
                for (int t = 0; t < m; ++t) {
                    double r2 = 0.0;
#pragma unroll
                    for (int d = 0; d < D; ++d) { const double e = q[d] - tile[d][t]; r2 = fma(e, e, r2); }
                    sum += exp(-0.5 * (r2 - d_min));
                }
            }
        }
    }
    if (j < n_queries) out[j] = log(sum) - 0.5 * d_min + log_norm;
}


// ---- ensemble-MCMC moves (emcee's red-blue scheme: the walkers of one half propose from the other half) ----------------
// One thread per moved walker.  Walkers are addressed through index lists (`mine`, `others`: the random split of the
// iteration), so the ensemble itself never moves; the gathers touch n_dim columns of a few MB that live in L2.
struct WalkerRng {
    Philox4x32 g;
    uint32_t w[4];
    int used;
    __device__ void init(uint64_t seed, uint64_t walker, uint32_t purpose) { g.init(seed, walker, purpose); used = 4; }
    __device__ uint32_t word() {
        if (used == 4) { g.next4(w); used = 0; }
        return w[used++];
    }
    __device__ double uniform() { const uint32_t hi = word(); return u53(hi, word()); }          // [0, 1)
    __device__ int64_t below(int64_t n) {                                                         // uniform on 0 .. n-1
        const uint32_t hi = word();
        return (int64_t)__umul64hi(((uint64_t)hi << 32) | word(), (uint64_t)n);
    }
    __device__ double normal() {                                                                  // Box-Muller, one of the pair
        const double u1 = 1.0 - uniform(), u2 = uniform();
        return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    }
};

__global__ void __launch_bounds__(128) walkers_propose_kernel(sodm_walker_move_args a) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.n_mine) return;
    const int D = a.n_dim;
    const int64_t me = a.mine[j];
    WalkerRng rng;
    rng.init(a.seed, (uint64_t)me, (uint32_t)a.iteration * 2u);
    const double* x = a.x;
    double lf = 0.0;
    if (a.kind == SODM_MOVE_STRETCH) {
        // Goodman & Weare: z ~ g(z) on [1/a, a], q = c + z (s - c), factor z^(D-1)
        const double u = rng.uniform();
        const double t = (a.param - 1.0) * u + 1.0;
        const double z = t * t / a.param;
        const int64_t c = a.others[rng.below(a.n_others)];
        for (int d = 0; d < D; ++d) {
            const double cd = x[d * a.ld + c];
            a.q[d * a.ld_q + j] = cd - (cd - x[d * a.ld + me]) * z;
        }
        lf = (D - 1.0) * log(z);
    } else if (a.kind == SODM_MOVE_DE) {
        // ter Braak: q = s + gamma (c_a - c_b) + eps, two distinct walkers of the other half, eps ~ N(0, sigma^2)
        const double gamma = a.param > 0.0 ? a.param : 2.38 / sqrt(2.0 * D);
        int64_t ia = rng.below(a.n_others), ib = rng.below(a.n_others - 1);
        ib += ib >= ia;
        const int64_t ca = a.others[ia], cb = a.others[ib];
        const double eps = a.sigma * rng.normal();
        for (int d = 0; d < D; ++d)
            a.q[d * a.ld_q + j] = x[d * a.ld + me] + gamma * (x[d * a.ld + ca] - x[d * a.ld + cb]) + eps;
    } else {
        // ter Braak & Vrugt snooker: along the line through s and z, a jump of gamma times the difference of the
        // projections of two more walkers; factor (|q - z| / |s - z|)^(D-1)
        int64_t i0 = rng.below(a.n_others), i1 = rng.below(a.n_others - 1), i2 = rng.below(a.n_others - 2);
        i1 += i1 >= i0;
        const int64_t lo = i0 < i1 ? i0 : i1, hi = i0 < i1 ? i1 : i0;
        i2 += i2 >= lo;
        i2 += i2 >= hi;
        const int64_t z = a.others[i0], z1 = a.others[i1], z2 = a.others[i2];
        double norm2 = 0.0, proj = 0.0;
        for (int d = 0; d < D; ++d) {
            const double delta = x[d * a.ld + me] - x[d * a.ld + z];
            norm2 = fma(delta, delta, norm2);
            proj = fma(delta, x[d * a.ld + z1] - x[d * a.ld + z2], proj);
        }
        const double norm = sqrt(norm2);
        const double step = a.param * proj / norm2;                  // gamma (u.z1 - u.z2) u = step * delta
        double qn2 = 0.0;
        for (int d = 0; d < D; ++d) {
            const double sd = x[d * a.ld + me], zd = x[d * a.ld + z];
            const double qd = sd + step * (sd - zd);
            a.q[d * a.ld_q + j] = qd;
            qn2 = fma(qd - zd, qd - zd, qn2);
        }
        lf = (D - 1.0) * (0.5 * log(qn2) - log(norm));
    }
    a.log_factor[j] = lf;
}

// KDE move: an independence proposal from the Gaussian KDE of the walkers `others` (kernel covariance L L^T).  Also writes
// the whitened coordinates L^-1 (centres | current positions | proposals) that the pair sum of the Hastings factor reads.
struct KdeFactor {                                                   // packed lower triangles, (i, j) at i (i + 1) / 2 + j
    double L[SODM_POPULATION_MAX_DIM * (SODM_POPULATION_MAX_DIM + 1) / 2];
    double Linv[SODM_POPULATION_MAX_DIM * (SODM_POPULATION_MAX_DIM + 1) / 2];
};

__global__ void __launch_bounds__(128) walkers_kde_kernel(sodm_walker_move_args a, KdeFactor f) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int D = a.n_dim;
    const int64_t K = a.n_others;
    double v[SODM_POPULATION_MAX_DIM], z[SODM_POPULATION_MAX_DIM];
    if (j < K) {                                                     // whitened kernel centres
        const int64_t c = a.others[j];
#pragma unroll
        for (int d = 0; d < SODM_POPULATION_MAX_DIM; ++d) v[d] = d < D ? a.x[d * a.ld + c] : 0.0;
#pragma unroll
        for (int i = 0; i < SODM_POPULATION_MAX_DIM; ++i) {
            if (i >= D) break;
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k <= i; ++k) acc = fma(f.Linv[i * (i + 1) / 2 + k], v[k], acc);
            a.white[i * a.ld_white + j] = acc;
        }
    }
    if (j >= a.n_mine) return;
    const int64_t me = a.mine[j];
    WalkerRng rng;
    rng.init(a.seed, (uint64_t)me, (uint32_t)a.iteration * 2u);
    const int64_t c = a.others[rng.below(K)];
#pragma unroll
    for (int d = 0; d < SODM_POPULATION_MAX_DIM; ++d) z[d] = d < D ? rng.normal() : 0.0;
    // proposal q = c + L z; whitened proposal = L^-1 c + z; whitened current position = L^-1 s
#pragma unroll
    for (int d = 0; d < SODM_POPULATION_MAX_DIM; ++d) v[d] = d < D ? a.x[d * a.ld + c] : 0.0;
#pragma unroll
    for (int i = 0; i < SODM_POPULATION_MAX_DIM; ++i) {
        if (i >= D) break;
        double q = v[i], w = 0.0;
#pragma unroll
        for (int k = 0; k <= i; ++k) {
            q = fma(f.L[i * (i + 1) / 2 + k], z[k], q);
            w = fma(f.Linv[i * (i + 1) / 2 + k], v[k], w);
        }
        a.q[i * a.ld_q + j] = q;
        a.white[i * a.ld_white + K + a.n_mine + j] = w + z[i];
    }
#pragma unroll
    for (int d = 0; d < SODM_POPULATION_MAX_DIM; ++d) v[d] = d < D ? a.x[d * a.ld + me] : 0.0;
#pragma unroll
    for (int i = 0; i < SODM_POPULATION_MAX_DIM; ++i) {
        if (i >= D) break;
        double w = 0.0;
#pragma unroll
        for (int k = 0; k <= i; ++k) w = fma(f.Linv[i * (i + 1) / 2 + k], v[k], w);
        a.white[i * a.ld_white + K + j] = w;
    }
}

// Clip a population into the bounds IN PLACE and note who had to be clipped: the reference's objective clips an
// out-of-bounds theta and scores it -inf (objective_functions.py:301-307).
__global__ void __launch_bounds__(256) population_clip_kernel(int n_dim, int64_t n, int64_t ld, double* __restrict__ x, Bounds b,
                                                              double* __restrict__ penalty) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        bool out = false;
        for (int d = 0; d < n_dim; ++d) {
            const double v = x[d * ld + i];
            if (v > b.ub[d]) { x[d * ld + i] = b.ub[d]; out = true; }
            else if (v < b.lb[d]) { x[d * ld + i] = b.lb[d]; out = true; }
        }
        penalty[i] = out ? -1.0 / 0.0 : 0.0;
    }
}

// Metropolis-Hastings: walker mine[j] moves to q[:, j] when log_factor + new_lp (+ penalty) - lp > log u.
__global__ void __launch_bounds__(128) walkers_accept_kernel(int n_dim, int64_t ld, double* __restrict__ x, double* __restrict__ lp,
                                                             const int64_t* __restrict__ mine, int64_t n_mine,
                                                             const double* __restrict__ q, int64_t ld_q,
                                                             const double* __restrict__ new_lp, const double* __restrict__ penalty,
                                                             const double* __restrict__ log_factor, uint64_t seed,
                                                             uint32_t iteration, int64_t* __restrict__ accepted,
                                                             int32_t* __restrict__ status) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_mine) return;
    const int64_t me = mine[j];
    WalkerRng rng;
    rng.init(seed, (uint64_t)me, iteration * 2u + 1u);
    const double cand = new_lp[j];
    if (cand != cand) {
        if (status) atomicOr(status, SODM_WALKERS_ST_NAN);
        return;
    }
    if (penalty && penalty[j] < 0.0) return;                         // clipped into the bounds: scored -inf
    const double log_u = log(1.0 - rng.uniform());                   // u on (0, 1]
    if (isinf(cand) || !(log_factor[j] + cand - lp[me] > log_u)) return;     // a non-finite score never moves a walker
    lp[me] = cand;
    for (int d = 0; d < n_dim; ++d) x[d * ld + me] = q[d * ld_q + j];
    accepted[me] += 1;
}

template <int D>
static int launch_kde(const double* points, int64_t n_points, int64_t ld_p, const double* queries, int64_t n_queries,
                      int64_t ld_q, double log_norm, double* out, cudaStream_t s) {
    const unsigned grid = (unsigned)((n_queries + 127) / 128);
    kde_logpdf_kernel<D><<<grid, 128, 0, s>>>(points, n_points, ld_p, queries, n_queries, ld_q, log_norm, out);
    SODM_CUDA_CHECK(cudaGetLastError());
    return SODM_OK;
}

static unsigned grid_for(int64_t n) {
    int64_t blocks = (n + 255) / 256;
    return (unsigned)(blocks < 1 ? 1 : (blocks > 148 * 8 ? 148 * 8 : blocks));
}

static int check_population(const char* who, int32_t n_dim, int64_t n, int64_t ld) {
    if (n_dim < 1 || n_dim > SODM_POPULATION_MAX_DIM)
        return fail(SODM_EINVAL, "%s: n_dim %d outside 1..%d", who, n_dim, SODM_POPULATION_MAX_DIM);
    if (n < 0 || ld < n) return fail(SODM_EINVAL, "%s: n = %lld, ld = %lld", who, (long long)n, (long long)ld);
    return SODM_OK;
}

}  // namespace sodm

using namespace sodm;

extern "C" int sodm_pso_move(const sodm_pso_args* a) {
    if (!a) return fail(SODM_EINVAL, "sodm_pso_move: null args");
    int rc = check_population("sodm_pso_move", a->n_dim, a->n, a->ld);
    if (rc != SODM_OK) return rc;
    if (a->n == 0) return SODM_OK;
    if (!a->x || !a->v || !a->p || !a->g || !a->lb || !a->ub) return fail(SODM_EINVAL, "sodm_pso_move: null array");
    if (a->iteration < 0 || a->iteration > 0xFFFFFFFFLL) return fail(SODM_EINVAL, "sodm_pso_move: iteration out of range");
    Bounds b;
    for (int d = 0; d < a->n_dim; ++d) {
        if (!(a->ub[d] > a->lb[d])) return fail(SODM_EINVAL, "sodm_pso_move: upper bound %d not above lower bound", d);
        b.lb[d] = a->lb[d];
        b.ub[d] = a->ub[d];
    }
    pso_move_kernel<<<grid_for(a->n), 256, 0, (cudaStream_t)a->stream>>>(a->n_dim, a->n, a->ld, a->x, a->v, a->p, a->g, b,
                                                                         a->omega, a->phip, a->phig, a->seed,
                                                                         (uint32_t)a->iteration, a->member_id0);
    SODM_CUDA_CHECK(cudaGetLastError());
    return SODM_OK;
}

extern "C" int sodm_pso_best(int32_t n_dim, int64_t n, int64_t ld, const double* x, const double* fx,
                             const uint8_t* feasible, double* p, double* fp, double* best, int64_t* best_index, void* stream) {
    int rc = check_population("sodm_pso_best", n_dim, n, ld);
    if (rc != SODM_OK) return rc;
    if (n == 0) return fail(SODM_EINVAL, "sodm_pso_best: empty swarm");
    if (!x || !fx || !p || !fp || !best || !best_index) return fail(SODM_EINVAL, "sodm_pso_best: null array");
    cudaStream_t s = (cudaStream_t)stream;
    pso_keep_kernel<<<grid_for(n), 256, 0, s>>>(n_dim, n, ld, x, fx, feasible, p, fp);
    SODM_CUDA_CHECK(cudaGetLastError());
    pso_argmin_kernel<<<1, 1024, 0, s>>>(n_dim, n, ld, p, fp, best, best_index);
    SODM_CUDA_CHECK(cudaGetLastError());
    return SODM_OK;
}

extern "C" int sodm_kde_logpdf(int32_t n_dim, const double* points, int64_t n_points, int64_t ld_points,
                               const double* queries, int64_t n_queries, int64_t ld_queries, double log_norm, double* out,
                               void* stream) {
    int rc = check_population("sodm_kde_logpdf (points)", n_dim, n_points, ld_points);
    if (rc == SODM_OK) rc = check_population("sodm_kde_logpdf (queries)", n_dim, n_queries, ld_queries);
    if (rc != SODM_OK) return rc;
    if (n_queries == 0) return SODM_OK;
    if (n_points == 0) return fail(SODM_EINVAL, "sodm_kde_logpdf: no kernels");
    if (!points || !queries || !out) return fail(SODM_EINVAL, "sodm_kde_logpdf: null array");
    cudaStream_t s = (cudaStream_t)stream;
    switch (n_dim) {
#define SODM_KDE_CASE(D) \
        case D: return launch_kde<D>(points, n_points, ld_points, queries, n_queries, ld_queries, log_norm, out, s);
        SODM_KDE_CASE(1) SODM_KDE_CASE(2) SODM_KDE_CASE(3) SODM_KDE_CASE(4) SODM_KDE_CASE(5) SODM_KDE_CASE(6)
        SODM_KDE_CASE(7) SODM_KDE_CASE(8) SODM_KDE_CASE(9) SODM_KDE_CASE(10) SODM_KDE_CASE(11) SODM_KDE_CASE(12)
        SODM_KDE_CASE(13) SODM_KDE_CASE(14) SODM_KDE_CASE(15) SODM_KDE_CASE(16)
#undef SODM_KDE_CASE
    }
    return fail(SODM_EINVAL, "sodm_kde_logpdf: n_dim %d", n_dim);
}

extern "C" int sodm_walkers_propose(const sodm_walker_move_args* a) {
    if (!a) return fail(SODM_EINVAL, "sodm_walkers_propose: null args");
    int rc = check_population("sodm_walkers_propose", a->n_dim, a->n_walkers, a->ld);
    if (rc != SODM_OK) return rc;
    if (a->n_mine < 0 || a->ld_q < a->n_mine) return fail(SODM_EINVAL, "sodm_walkers_propose: n_mine = %lld, ld_q = %lld", (long long)a->n_mine, (long long)a->ld_q);
    if (a->n_mine == 0) return SODM_OK;
    const int need = a->kind == SODM_MOVE_STRETCH ? 1 : a->kind == SODM_MOVE_DE ? 2 : a->kind == SODM_MOVE_SNOOKER ? 3 :
                     a->kind == SODM_MOVE_KDE ? 1 : -1;
    if (need < 0) return fail(SODM_EINVAL, "sodm_walkers_propose: unknown move %d", a->kind);
    if (a->n_others < need) return fail(SODM_EINVAL, "sodm_walkers_propose: the move needs %d walkers in the other half, got %lld", need, (long long)a->n_others);
    if (a->kind == SODM_MOVE_STRETCH && !(a->param > 1.0)) return fail(SODM_EINVAL, "sodm_walkers_propose: stretch scale a must exceed 1");
    if (!a->x || !a->mine || !a->others || !a->q) return fail(SODM_EINVAL, "sodm_walkers_propose: null array");
    if (a->iteration < 0 || a->iteration > 0x7FFFFFFFLL) return fail(SODM_EINVAL, "sodm_walkers_propose: iteration out of range");
    if (a->kind == SODM_MOVE_KDE) {
        if (!a->chol || !a->white) return fail(SODM_EINVAL, "sodm_walkers_propose: the KDE move needs chol and white");
        if (a->ld_white < a->n_others + 2 * a->n_mine) return fail(SODM_EINVAL, "sodm_walkers_propose: ld_white = %lld < n_others + 2 n_mine", (long long)a->ld_white);
        const int D = a->n_dim;
        KdeFactor f = {};
        for (int i = 0; i < D; ++i) {
            if (!(a->chol[i * D + i] > 0.0)) return fail(SODM_EINVAL, "sodm_walkers_propose: chol[%d][%d] is not positive", i, i);
            for (int j = 0; j <= i; ++j) f.L[i * (i + 1) / 2 + j] = a->chol[i * D + j];
        }
        for (int j = 0; j < D; ++j)                                  // L^-1 by forward substitution, column by column
            for (int i = j; i < D; ++i) {
                double acc = i == j ? 1.0 : 0.0;
                for (int k = j; k < i; ++k) acc -= f.L[i * (i + 1) / 2 + k] * f.Linv[k * (k + 1) / 2 + j];
                f.Linv[i * (i + 1) / 2 + j] = acc / f.L[i * (i + 1) / 2 + i];
            }
        const int64_t n = a->n_mine > a->n_others ? a->n_mine : a->n_others;
        walkers_kde_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)a->stream>>>(*a, f);
        SODM_CUDA_CHECK(cudaGetLastError());
        return SODM_OK;
    }
    if (!a->log_factor) return fail(SODM_EINVAL, "sodm_walkers_propose: null array");
    walkers_propose_kernel<<<(unsigned)((a->n_mine + 127) / 128), 128, 0, (cudaStream_t)a->stream>>>(*a);
    SODM_CUDA_CHECK(cudaGetLastError());
    return SODM_OK;
}

extern "C" int sodm_population_clip(int32_t n_dim, int64_t n, int64_t ld, double* x, const double* lb, const double* ub,
                                    double* penalty, void* stream) {
    int rc = check_population("sodm_population_clip", n_dim, n, ld);
    if (rc != SODM_OK) return rc;
    if (n == 0) return SODM_OK;
    if (!x || !lb || !ub || !penalty) return fail(SODM_EINVAL, "sodm_population_clip: null array");
    Bounds b;
    for (int d = 0; d < n_dim; ++d) { b.lb[d] = lb[d]; b.ub[d] = ub[d]; }
    population_clip_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(n_dim, n, ld, x, b, penalty);
    SODM_CUDA_CHECK(cudaGetLastError());
    return SODM_OK;
}

extern "C" int sodm_walkers_accept(int32_t n_dim, int64_t n_walkers, int64_t ld, double* x, double* lp, const int64_t* mine,
                                   int64_t n_mine, const double* q, int64_t ld_q, const double* new_lp, const double* penalty,
                                   const double* log_factor, uint64_t seed, int64_t iteration, int64_t* accepted,
                                   int32_t* status, void* stream) {
    int rc = check_population("sodm_walkers_accept", n_dim, n_walkers, ld);
    if (rc != SODM_OK) return rc;
    if (n_mine < 0 || ld_q < n_mine) return fail(SODM_EINVAL, "sodm_walkers_accept: n_mine = %lld, ld_q = %lld", (long long)n_mine, (long long)ld_q);
    if (n_mine == 0) return SODM_OK;
    if (!x || !lp || !mine || !q || !new_lp || !log_factor || !accepted) return fail(SODM_EINVAL, "sodm_walkers_accept: null array");
    if (iteration < 0 || iteration > 0x7FFFFFFFLL) return fail(SODM_EINVAL, "sodm_walkers_accept: iteration out of range");
    walkers_accept_kernel<<<(unsigned)((n_mine + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        n_dim, ld, x, lp, mine, n_mine, q, ld_q, new_lp, penalty, log_factor, seed, (uint32_t)iteration, accepted, status);
    SODM_CUDA_CHECK(cudaGetLastError());
    return SODM_OK;
}
