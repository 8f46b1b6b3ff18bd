// Observation-noise samplers for ensemble post-processing: Poisson, Gaussian and negative binomial variates drawn
// per (cell, trajectory) element from a Philox stream keyed by (seed, global trajectory id, cell).
//
// Replaces /root/reference/src/pySODM/optimization/utils.py:11-95 (`add_poisson_noise`, `add_gaussian_noise`,
// `add_negative_binomial_noise`), which call NumPy's legacy global generator on the whole ensemble on the host:
//   np.random.poisson(x)                                     utils.py:30
//   np.random.normal(x, sigma) / np.random.normal(x, |sigma x|)   utils.py:62,64
//   np.random.negative_binomial(1/alpha, (1/alpha)/(x + 1/alpha))  utils.py:91-92
// NumPy's legacy samplers are not in the reference tree (numpy is an unpinned dependency, setup.py:19; 2.3 here):
// `legacy_random_poisson` = multiplication method for lam < 10, Hoermann's PTRS otherwise; `legacy_negative_binomial`
// = Poisson(Gamma(n, (1-p)/p)); `legacy_gauss` = polar Box-Muller.  Parity is in LAW (KS against NumPy draws).
#pragma once
#include <cmath>
#include <cstdint>
#include "philox.cuh"

namespace sodm {

// Per-element random stream: blocks of four words on demand.
struct ElementRng {
    Philox4x32 gen;
    uint32_t w[4];
    int used;
    __host__ __device__ __forceinline__ void init(uint64_t seed, uint64_t traj, uint32_t cell) {
        gen.init(seed ^ 0x6e6f697365ull /* "noise" */, traj, cell);
        used = 4;
    }
    __host__ __device__ __forceinline__ uint32_t next() {
        if (used == 4) { gen.next4(w); used = 0; }
        return w[used++];
    }
    // uniform in (0, 1) with 53 random bits: never 0, never 1
    __host__ __device__ __forceinline__ double uniform() {
        const uint64_t hi = next(), lo = next();
        return ((double)(((hi << 32) | lo) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    }
    __host__ __device__ __forceinline__ double normal() {                 // Box-Muller, one of the pair
        const double r = sqrt(-2.0 * log(uniform()));
        const double phi = 6.283185307179586 * ((double)next() + 0.5) * (1.0 / 4294967296.0);
        return r * cos(phi);
    }
};

// Poisson(lam), lam >= 0 finite.  lam < 10: sequential-search inversion; else Hoermann's PTRS ("The transformed
// rejection method for generating Poisson random variables", 1993) with the exact test in FP64.
__host__ __device__ inline double sample_poisson(ElementRng& rng, double lam) {
    if (lam == 0.0) return 0.0;
    if (lam < 10.0) {
        double f = exp(-lam), u = rng.uniform(), k = 0.0;
        while (u > f && k < 1024.0) {
            u -= f;
            k += 1.0;
            f *= lam / k;
        }
        return k;
    }
    const double slam = sqrt(lam), loglam = log(lam);
    const double b = 0.931 + 2.53 * slam;
    const double a = -0.059 + 0.02483 * b;
    const double inv_alpha = 1.1239 + 1.1328 / (b - 3.4);
    const double v_r = 0.9277 - 3.6224 / (b - 2.0);
    for (int attempt = 0; attempt < 1000; ++attempt) {
        const double u = rng.uniform() - 0.5, v = rng.uniform();
        const double us = 0.5 - fabs(u);
        const double k = floor((2.0 * a / us + b) * u + lam + 0.43);
        if (us >= 0.07 && v <= v_r) return k;
        if (k < 0.0 || (us < 0.013 && v > us)) continue;
        if (log(v) + log(inv_alpha) - log(a / (us * us) + b) <= -lam + k * loglam - lgamma(k + 1.0)) return k;
    }
    return floor(lam);        // unreachable in practice (acceptance > 0.86 per attempt)
}

// Gamma(shape, 1): Marsaglia & Tsang (2000) for shape >= 1, boosted by U^(1/shape) below.
__host__ __device__ inline double sample_standard_gamma(ElementRng& rng, double shape) {
    const double boost = shape < 1.0 ? pow(rng.uniform(), 1.0 / shape) : 1.0;
    const double d = (shape < 1.0 ? shape + 1.0 : shape) - 1.0 / 3.0;
    const double c = 1.0 / sqrt(9.0 * d);
    for (int attempt = 0; attempt < 1000; ++attempt) {
        const double z = rng.normal();
        const double t = 1.0 + c * z;
        if (t <= 0.0) continue;
        const double v = t * t * t;
        const double u = rng.uniform();
        if (u < 1.0 - 0.0331 * (z * z) * (z * z)) return d * v * boost;
        if (log(u) < 0.5 * z * z + d * (1.0 - v + log(v))) return d * v * boost;
    }
    return d * boost;
}

enum NoiseStatus : int32_t { NOISE_OK = 0, NOISE_BAD_MEAN = 1, NOISE_BAD_PARAM = 2 };

// One noisy value for model output x.  kind / param as in sodm_noise_args.
__host__ __device__ inline double noisy_value(ElementRng& rng, int kind, double param, double x, int32_t& status) {
    switch (kind) {
        case 0:                                                   // SODM_NOISE_POISSON
            if (!(x >= 0.0) || isinf(x)) { status |= NOISE_BAD_MEAN; return x; }       // NumPy: "lam < 0 or lam contains NaNs"
            return sample_poisson(rng, x);
        case 1:                                                   // SODM_NOISE_GAUSSIAN_ABS
            return x + param * rng.normal();
        case 2:                                                   // SODM_NOISE_GAUSSIAN_REL
            return x + fabs(param * x) * rng.normal();
        default: {                                                // SODM_NOISE_NEGBIN
            // n = 1/alpha, p = n / (x + n); NumPy raises for p outside [0, 1] or NaN, i.e. for x < 0 or NaN
            if (!(x >= 0.0) || isinf(x)) { status |= NOISE_BAD_MEAN; return x; }
            const double n = 1.0 / param;
            if (x == 0.0) return 0.0;                             // p = 1: Gamma scale 0
            return sample_poisson(rng, sample_standard_gamma(rng, n) * (x / n));
        }
    }
}

}  // namespace sodm
