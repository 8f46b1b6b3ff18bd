// Observation-noise samplers for ensemble post-processing: Poisson, Gaussian and negative binomial variates drawn
// per (cell, trajectory) element from Philox streams keyed by (seed, global trajectory id, cell).
//
// Replaces /root/reference/src/pySODM/optimization/utils.py:11-95 (`add_poisson_noise`, `add_gaussian_noise`,
// `add_negative_binomial_noise`), which call NumPy's legacy global generator on the whole ensemble on the host:
//   np.random.poisson(x)                                     utils.py:30
//   np.random.normal(x, sigma) / np.random.normal(x, |sigma x|)   utils.py:62,64
//   np.random.negative_binomial(1/alpha, (1/alpha)/(x + 1/alpha))  utils.py:91-92
// NumPy's legacy samplers are not in the reference tree (numpy is an unpinned dependency, setup.py:19; 2.3 here):
// `legacy_random_poisson` = multiplication method for lam < 10, Hoermann's PTRS otherwise; `legacy_negative_binomial`
// = Poisson(Gamma(n, (1-p)/p)); `legacy_gauss` = polar Box-Muller.  Parity is in LAW (KS against NumPy draws).
//
// Precision.  The samplers run in FP32 except where a quantity of the size of the MEAN enters (candidate k and
// k - lam, kept in FP64 so that means of 10^7 lose nothing).  What makes FP32 sufficient for the acceptance test is its
// FORM:   log f(k) = k log(lam) - lam - log k!   =   -D(k, lam) - delta(k) - 1/2 log(2 pi k)
// with the deviance D(k, lam) = k log(k / lam) + lam - k >= 0 and delta(k) = log k! - log(sqrt(2 pi k) (k/e)^k)
// (Loader 2000, "Fast and accurate computation of binomial probabilities"): every term is of the size of the result,
// nothing cancels, and log f comes out to ~1e-6 absolute at any mean -- no FP64 log / lgamma anywhere.
#pragma once
#include <cmath>
#include <cstdint>
#include "philox.cuh"
#include "binomial.cuh"      // fast_div / fast_log / fast_exp / fast_sqrt (approximate on the device, exact on the host)

namespace sodm {

constexpr uint64_t NOISE_SEED_SALT = 0x6e6f697365ull;   // "noise"

// Per-element random stream, handed out as PAIRS of words: a Philox block is two pairs, and which pair comes next is
// one predicate -- no indexing into the block (a dynamically indexed register array costs a select chain per word).
struct ElementRng {
    Philox4x32 gen;
    uint32_t w0, w1, w2, w3;
    bool second;
    __host__ __device__ __forceinline__ void init(uint64_t seed, uint64_t traj, uint32_t cell) {
        gen.init(seed ^ NOISE_SEED_SALT, traj, cell);
        second = false;
    }
    __host__ __device__ __forceinline__ void pair(uint32_t& a, uint32_t& b) {
        if (!second) {
            uint32_t w[4];
            gen.next4(w);
            w0 = w[0]; w1 = w[1]; w2 = w[2]; w3 = w[3];
        }
        a = second ? w2 : w0;
        b = second ? w3 : w1;
        second = !second;
    }
};

__host__ __device__ __forceinline__ float u01f(uint32_t w) { return u01(w); }   // [2^-26, 1 - 2^-24], 24 bits (philox.cuh)

// Two standard normal variates from two words (Box-Muller, both outputs).  The radius takes 32 bits of `wr`, the angle
// 24 bits of `wa`; should `wr` be 0 the tail would stop at 6.7 sigma, so the low bits of `wa` extend it (to 8.2 sigma).
__host__ __device__ __forceinline__ void normal_pair(uint32_t wr, uint32_t wa, float& z0, float& z1) {
    const float u = wr ? ((float)wr + 0.5f) * 2.3283064365386963e-10f : ((float)(wa & 0xffffu) + 0.5f) * 3.552713678800501e-15f;
    // hardware log2 / rsqrt / sin / cos on the device: absolute errors of ~2^-21 in the log and in the angle functions
    // move a variate by ~1e-6 of a standard deviation, far below anything a noisy observation resolves
    const float r = fast_sqrt(-2.0f * fast_log(fminf(u, 0.99999994f)));
    const float phi = 6.2831853f * ((float)(wa >> 8) * 5.9604644775390625e-8f - 0.5f);       // angle in [-pi, pi)
    float s, c;
#ifdef __CUDA_ARCH__
    __sincosf(phi, &s, &c);
#else
    s = sinf(phi);
    c = cosf(phi);
#endif
    z0 = r * c;
    z1 = r * s;
}
// Four normal variates for four consecutive cells of one trajectory from ONE Philox block: the Gaussian-noise kernel's
// generator cost per element is a quarter of a block.  Stream = (seed, trajectory, group of four cells).
__host__ __device__ __forceinline__ void normal_quad(uint64_t seed, uint64_t traj, uint32_t cell_group, float (&z)[4]) {
    Philox4x32 gen;
    gen.init(seed ^ NOISE_SEED_SALT, traj, cell_group | 0x80000000u);
    uint32_t w[4];
    gen.next4(w);
    normal_pair(w[0], w[1], z[0], z[1]);
    normal_pair(w[2], w[3], z[2], z[3]);
}

__host__ __device__ __forceinline__ float stirling_delta_f(float x) {        // x >= 1 integer-valued; 4e-6 at x = 2
    const float r = fast_div(1.0f, x);
    const float r2 = r * r;
    const float series = (0.0833333333f - (0.00277777778f - 0.000793650794f * r2) * r2) * r;
    return x < 1.5f ? 0.0810614668f : series;
}
// D(k, lam) = k log(k / lam) + lam - k for k >= 1, lam > 0, from d = k - lam given EXACTLY (FP64 at large means):
// D = d v + 2 k v (v^2/3 + v^4/5 + v^6/7 + v^8/9 + ...), v = d / (k + lam); first dropped term 5e-8 relative at |v| = 0.2
__host__ __device__ __forceinline__ float deviance_f(float k, float lam, float d) {
    const float sum = k + lam;
    if (fabsf(d) < 0.2f * sum) {
        const float v = fast_div(d, sum);
        const float v2 = v * v;
        return fmaf(2.0f * k * v * v2, fmaf(v2, fmaf(v2, fmaf(v2, 1.0f / 9.0f, 1.0f / 7.0f), 0.2f), 1.0f / 3.0f), d * v);
    }
    return k * logf(k / lam) - d;
}

// Poisson(lam), lam >= 0 finite.  lam < 10: sequential-search inversion; else Hoermann's PTRS ("The transformed
// rejection method for generating Poisson random variables", 1993).
// `seeded`: the first attempt uses the given words (the negative-binomial path hands over the unused bits of the block
// its gamma variate came from, so that a typical element costs ONE Philox block); later attempts draw fresh pairs.
__host__ __device__ inline double sample_poisson(ElementRng& rng, double lam, bool seeded = false, uint32_t wu0 = 0u, uint32_t wv0 = 0u) {
    if (lam == 0.0) return 0.0;
    const float lf = (float)lam;
    if (lam < 10.0) {
        uint32_t wa = wu0, wb;
        if (!seeded) rng.pair(wa, wb);
        float f = fast_exp(-lf), u = u01f(wa), k = 0.0f;
        // the search ends at the latest where the pmf has dropped below the 2^-24 grid of u (far past the mode): round-off
        // in the running difference can then no longer carry it to the cap (see binomial_inversion_word)
        while (u > f && f > 1e-8f && k < 1024.0f) {
            u -= f;
            k += 1.0f;
            f *= fast_div(lf, k);
        }
        return (double)k;
    }
    // hat constants in FP32 with approximate quotients: any consistent set defines a valid hat, their rounding only
    // moves the proposal by ~1e-7
    const float slam = fast_sqrt(lf);
    const float b = 0.931f + 2.53f * slam;
    const float a = -0.059f + 0.02483f * b;
    const float inv_alpha = 1.1239f + fast_div(1.1328f, b - 3.4f);
    const float v_r = 0.9277f - fast_div(3.6224f, b - 2.0f);
    const bool wide = lam >= 4194304.0;                                      // FP32 no longer resolves k - lam
    for (int attempt = 0; attempt < 1000; ++attempt) {
        uint32_t wu = wu0, wv = wv0;
        if (!seeded || attempt > 0) rng.pair(wu, wv);
        const float u = u01f(wu) - 0.5f, v = u01f(wv);
        const float us = 0.5f - fabsf(u);
        const float t = (fast_div(2.0f * a, us) + b) * u;
        double k;
        float d;                                                             // k - lam
        if (wide) {
            k = floor((double)t + lam + 0.43);
            d = (float)(k - lam);
        } else {
            const float kf = floorf(t + lf + 0.43f);
            k = (double)kf;
            d = kf - lf;
        }
        if (us >= 0.07f && v <= v_r) return k;
        if (k < 0.0 || (us < 0.013f && v > us)) continue;
        const float lhs = fast_log(fast_div(v * inv_alpha, fast_div(a, us * us) + b));
        const float kf = (float)k;
        const float rhs = k < 0.5 ? -lf : -deviance_f(kf, lf, d) - stirling_delta_f(kf) - 0.5f * fast_log(6.2831853f * kf);
        if (lhs <= rhs) return k;
    }
    return floor(lam);        // unreachable in practice (acceptance > 0.86 per attempt)
}

// Gamma(shape, 1): Marsaglia & Tsang (2000) for shape >= 1, boosted by U^(1/shape) below.  One Philox block per attempt:
// (wr, wa) -> the normal variate, the top 24 bits of wu -> the acceptance uniform, wb -> the boost uniform (shape < 1).
// `spare_u`, `spare_v` return words for ONE further attempt of a downstream sampler, built from bits this attempt did
// not look at: wb as it is (shape >= 1), and the low bytes of wa, wu, wb (each consumer reads the top 24 bits of a word).
__host__ __device__ inline float sample_standard_gamma(ElementRng& rng, float shape, bool& spare, uint32_t& spare_u, uint32_t& spare_v) {
    const float d = (shape < 1.0f ? shape + 1.0f : shape) - 1.0f / 3.0f;
    const float c = fast_div(1.0f, fast_sqrt(9.0f * d));
    for (int attempt = 0; attempt < 1000; ++attempt) {
        uint32_t wr, wa, wu, wb;
        rng.pair(wr, wa);
        rng.pair(wu, wb);
        float z, z_unused;
        normal_pair(wr, wa, z, z_unused);
        const float t = 1.0f + c * z;
        if (t <= 0.0f) continue;
        const float v = t * t * t;
        const float u = u01f(wu);
        // second test only for the ~3 % the squeeze leaves; the hardware logarithm's absolute error (2^-21) times d moves
        // this acceptance threshold by ~1e-5
        const bool ok = u < 1.0f - 0.0331f * (z * z) * (z * z) || fast_log(u) < 0.5f * z * z + d * (1.0f - v + fast_log(v));
        if (ok) {
            spare = shape >= 1.0f;
            spare_u = wb;
            spare_v = ((wa & 0xffu) << 24) | ((wu & 0xffu) << 16) | ((wb & 0xffu) << 8);
            return d * v * (shape < 1.0f ? powf(u01f(wb), 1.0f / shape) : 1.0f);
        }
    }
    spare = false;
    return d;
}

enum NoiseStatus : int32_t { NOISE_OK = 0, NOISE_BAD_MEAN = 1 };

// Count noise around model output x: Poisson (NEGBIN = false) or negative binomial with n = 1 / alpha.
template <bool NEGBIN>
__host__ __device__ __forceinline__ double count_noise(ElementRng& rng, double alpha, double x, int32_t& status) {
    // NumPy raises for a negative / NaN mean ("lam < 0 or lam contains NaNs"; p = n / (x + n) outside [0, 1])
    if (!(x >= 0.0) || isinf(x)) { status |= NOISE_BAD_MEAN; return x; }
    if (x == 0.0) return 0.0;                                                // Poisson(0); NB with p = 1
    if (!NEGBIN) return sample_poisson(rng, x);
    const double n = 1.0 / alpha;
    bool spare;
    uint32_t su, sv;
    const float g = sample_standard_gamma(rng, (float)n, spare, su, sv);
    return sample_poisson(rng, (double)g * (x / n), spare, su, sv);
}

__host__ __device__ __forceinline__ double gaussian_noise(bool relative, double sigma, double x, float z) {
    return x + (relative ? fabs(sigma * x) : sigma) * (double)z;
}

}  // namespace sodm
