// Exact binomial sampling for one SIMT lane.
//
// The reference draws its transition counts with NumPy's legacy `RandomState.binomial`
// (/root/reference/src/pySODM/models/base.py:303): sequential inversion when n*p <= 30, BTPE
// otherwise.  Both are exact samplers, so any other exact sampler gives the same law; the ones
// below are chosen for SIMT execution:
//   * n*min(p,1-p) <  SODM_BINV_THRESHOLD : sequential inversion from 0 (short, FP32, one uniform);
//   * otherwise: Hoermann's BTRS transformed rejection (1993).  One attempt = two uniforms; ~80 % of
//     attempts are accepted by BTRS's own box test, which needs no logarithm.  The rest are settled
//     in FP32 by the quadratic squeeze of Kachitvichyanukul & Schmeiser's BTPE (step 5.2: for
//     x = |k - mode| < npq/2 - 1,  -x^2/(2npq) - rho <= log f(k)/f(mode) <= -x^2/(2npq) + rho), widened
//     by an FP32 round-off margin; only the sliver in between (~0.5-5 % of attempts) evaluates the
//     exact FP64 log-pmf ratio, whose Stirling terms cancel from O(n*p) to O(1) and need double.
// p > 0.5 is handled by the symmetry Bin(n, p) = n - Bin(n, 1 - p).
//
// `binomial_try` performs ONE attempt and reports whether it finished, so that the lanes of a warp
// can each work through their own list of draws without waiting for each other's rejections
// (see sample_worklist in jump_run.cuh).
#pragma once
#include <cmath>
#include <cstdint>
#include "philox.cuh"

namespace sodm {

#ifndef SODM_BINV_THRESHOLD
#define SODM_BINV_THRESHOLD 10.0f
#endif
// Not a free tuning knob: BTRS's hat dominates the pmf only from n p >= 10 on (Hoermann 1993), and the bounded inversion
// search below stops where the pmf has fallen under 1e-8, which P(X = 0) = q^n must therefore exceed: q^n >= 2^-(2 T) for
// n p < T, p <= 1/2.  (A tuning build with T = 20 returned 0 for every draw with q^n < 1e-8 and shifted the ensemble mean by 1 %.)
static_assert(SODM_BINV_THRESHOLD == 10.0f, "the inversion / BTRS hand-over is fixed by the validity ranges of the two samplers");

// Buffered 32-bit words from a Philox stream (4 words per block, no word is wasted).
struct PhiloxWords {
    Philox4x32 gen;
    uint32_t w0, w1, w2, w3;
    int pos;

    __host__ __device__ __forceinline__ void init(uint64_t seed, uint64_t traj, uint32_t purpose = 0) {
        gen.init(seed, traj, purpose);
        pos = 4;
        w0 = w1 = w2 = w3 = 0u;
    }
    __host__ __device__ __forceinline__ uint32_t next() {
        if (pos >= 4) {
            uint32_t b[4];
            gen.next4(b);
            w0 = b[0]; w1 = b[1]; w2 = b[2]; w3 = b[3];
            pos = 0;
        }
        const uint32_t v = pos == 0 ? w0 : (pos == 1 ? w1 : (pos == 2 ? w2 : w3));
        ++pos;
        return v;
    }
};

// ---- fast device arithmetic for the PROPOSAL side of the rejection sampler -------------------------------------
// The hat function / squeeze of BTRS may be evaluated with approximate division and logarithm: an error of a few ulp
// there moves a candidate boundary or an acceptance threshold by ~1e-7 relative, i.e. perturbs a uniform variate by
// 1e-7 -- invisible in the sampled law.  The exact pmf ratio (binomial_log_ratio*) keeps accurate arithmetic.
// On the device these are the bare special-function-unit instructions (MUFU.RCP / RSQ / EX2 / LG2 through the .ftz PTX
// forms): `__fdividef`, `rsqrtf`, `__expf` and `__logf` wrap the same instruction in a denormal-range rescue (5 / 6 / 4 / 4
// SASS instructions instead of 2) that cannot trigger here -- every operand is a normal float (us >= 2^-25, n p (1 - p) > 1,
// arguments of log in (1e-12, 1e12)) -- so the values are bit-identical and a fifth of the kernel's instructions goes away
// (ncu, profiles/r1_final_tau_leap_ncu_full.txt: 15.6 % of all warp instructions were __fdividef).
__host__ __device__ __forceinline__ float fast_div(float a, float b) {
#ifdef __CUDA_ARCH__
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    return r * a;
#else
    return a / b;
#endif
}
__host__ __device__ __forceinline__ float fast_sqrt(float x) {      // x > 0
#ifdef __CUDA_ARCH__
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return x * r;
#else
    return sqrtf(x);
#endif
}
__host__ __device__ __forceinline__ float fast_exp(float x) {
#ifdef __CUDA_ARCH__
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
    return r;
#else
    return expf(x);
#endif
}
__host__ __device__ __forceinline__ float fast_log(float x) {
#ifdef __CUDA_ARCH__
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r * 0.693147181f;
#else
    return logf(x);
#endif
}

#define SODM_STIRLING_TABLE {0.0810614667953272, 0.0413406959554092, 0.0276779256849983, 0.02079067210376509, \
                             0.0166446911898211, 0.0138761288230707, 0.0118967099458917, 0.0104112652619720,  \
                             0.00925546218271273, 0.00833056343336287}
#ifdef __CUDACC__
static __constant__ double kStirlingTailDev[10] = SODM_STIRLING_TABLE;
#endif

// log(k!) - [ (k+1/2) log(k+1) - (k+1) + 1/2 log(2 pi) ]   (Stirling tail fc(k), Hoermann 1993)
__host__ __device__ __forceinline__ double stirling_tail(double k) {
    if (k <= 9.0) {
#ifdef __CUDA_ARCH__
        return kStirlingTailDev[(int)k];
#else
        static const double tab[10] = SODM_STIRLING_TABLE;
        return tab[(int)k];
#endif
    }
    const double inv = 1.0 / (k + 1.0);                   // one FP64 division instead of three (each ~30 instructions)
    const double inv2 = inv * inv;
    return (1.0 / 12.0 - (1.0 / 360.0 - (1.0 / 1260.0) * inv2) * inv2) * inv;
}

// Sequential-search inversion; requires 0 < p <= 0.5 and n*p small.
template <class Rng>
__host__ __device__ __forceinline__ int binomial_inversion(Rng& rng, int n, float p) {
    const float q = 1.0f - p;
    const float s = p / q;
    float f = expf((float)n * log1pf(-p));      // q^n = P(X = 0)
    float u = u01(rng.next());
    int k = 0;
    // P(X=k+1) = P(X=k) * s * (n-k)/(k+1).  The cap only guards against FP32 round-off in the far tail.
    while (u > f && f > 1e-8f && k < n && k < 1024) {      // f > 1e-8: see binomial_inversion_word
        u -= f;
        ++k;
        f *= s * (float)(n - k + 1) / (float)k;
    }
    return k;
}

// exact log f(k)/f(m) for Bin(n, p), m = floor((n+1)p)   (Hoermann 1993, BTRS step 3.2).  Rarely executed (modes
// >= SODM_F32_MODE_LIMIT that the squeeze cannot settle) and large (three FP64 logarithms, twelve FP64 divisions):
// kept out of line so that it does not occupy the instruction cache of the hot loop.
#ifdef __CUDA_ARCH__
__device__ __noinline__
#else
inline
#endif
double binomial_log_ratio(double nd, double pd, double m, double kd) {
    const double r = pd / (1.0 - pd);
    return (m + 0.5) * log((m + 1.0) / (r * (nd - m + 1.0))) + stirling_tail(m) + stirling_tail(nd - m)
         + (nd + 1.0) * log((nd - m + 1.0) / (nd - kd + 1.0))
         + (kd + 0.5) * log(r * (nd - kd + 1.0) / (kd + 1.0))
         - stirling_tail(kd) - stirling_tail(nd - kd);
}

// ---- the same ratio for LARGE modes (>= SODM_F32_MODE_LIMIT) in FP64 multiply-adds only ---------------------------------
// With a = n - m + 1, b = n - k + 1, x = k - m and N1 = (m + 1) - p (n + 2) (a number in (-p, 1 - p], computed without
// cancellation trouble in FP64) the three logarithms of binomial_log_ratio are
//     log((m+1) / (r a)) = log1p(N1 / (p a)),    log(a / b) = log1p(x / b),    log(r b / (k+1)) = log1p(-(N1 + x) / (q (k+1))),
// all of arguments of size x / (n p q) <= 4 / sigma: log1p(N / D) = 2 atanh(N / (2 D + N)) needs ONE reciprocal (FP32 seed + two
// Newton steps) and four series terms.  The Stirling tails reduce to x / 12 (1 / ((m+1)(k+1)) - 1 / (a b)) (next terms < 1e-11).
// ~60 instructions instead of three FP64 logarithms and twelve FP64 divisions (~1000): the FP64 branch ran at 1.1 of 32 lanes
// and cost 4 % of the kernel (profiles/r1_mufu_tau_leap_ncu_full.txt).  Valid for |x| <= 0.05 min(b, q (k+1)); callers fall
// back to binomial_log_ratio outside (far tails, never seen in practice).
__host__ __device__ __forceinline__ double rcp_newton(double d) {          // d > 0, normal in FP32 range
#ifdef __CUDA_ARCH__
    float seed;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(seed) : "f"((float)d));
    double r = (double)seed;
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    return fma(r, e, r);
#else
    return 1.0 / d;
#endif
}
__host__ __device__ __forceinline__ double log1p_ratio(double num, double den) {      // log(1 + num / den), |num / den| <= 0.06
    const double s = num * rcp_newton(2.0 * den + num);
    const double s2 = s * s;
    return 2.0 * s * fma(s2, fma(s2, fma(s2, fma(s2, 1.0 / 9.0, 1.0 / 7.0), 0.2), 1.0 / 3.0), 1.0);
}
__host__ __device__ __forceinline__ bool binomial_log_ratio_large(double nd, double pd, double m, double kd, double& out) {
    const double x = kd - m, q = 1.0 - pd;
    const double a = nd - m + 1.0, b = nd - kd + 1.0;
    const double n1 = (m + 1.0) - pd * (nd + 2.0);
    const double d3 = q * (kd + 1.0);
    if (!(fabs(x) <= 0.05 * b && fabs(n1 + x) <= 0.05 * d3 && kd >= 0.0)) return false;
    const double tails = x * (1.0 / 12.0) * (rcp_newton((m + 1.0) * (kd + 1.0)) - rcp_newton(a * b));
    out = (m + 0.5) * log1p_ratio(n1, pd * a) + (nd + 1.0) * log1p_ratio(x, b) + (kd + 0.5) * log1p_ratio(-(n1 + x), d3) + tails;
    return true;
}

__host__ __device__ __forceinline__ float stirling_tail_f(float k) {
    // series in 1/(k+1); its error is 3e-6 at k = 1 and smaller beyond, so only k = 0 needs the tabulated value
    // (a per-lane table lookup would serialise in the constant cache)
    const float r1 = fast_div(1.0f, k + 1.0f);
    const float inv = r1 * r1;
    const float series = (0.0833333333f - (0.00277777778f - 0.000793650794f * inv) * inv) * r1;
    return k < 0.5f ? 0.0810614668f : series;
}
// the same for k >= 9 (the mode and n - mode of a BTRS variate: n p >= 10, p <= 1/2): two terms are exact to 1e-8
__host__ __device__ __forceinline__ float stirling_tail_large_f(float k) {
    const float r1 = fast_div(1.0f, k + 1.0f);
    return (0.0833333333f - 0.00277777778f * r1 * r1) * r1;
}

// log(1 + x) to FP32 accuracy RELATIVE to the result (what the (n + 1) log(...) term needs): for |x| < 0.4 the
// odd series of 2 atanh(x / (2 + x)) -- next term s^11 / 11 < 2e-9 relative -- else the library routine.
__host__ __device__ __forceinline__ float log1p_rel(float x) {
    if (fabsf(x) < 0.4f) {
        const float s = fast_div(x, 2.0f + x);
        const float s2 = s * s;
        return 2.0f * s * fmaf(s2, fmaf(s2, fmaf(s2, fmaf(s2, 1.0f / 9.0f, 1.0f / 7.0f), 0.2f), 1.0f / 3.0f), 1.0f);
    }
    return log1pf(x);
}

// The same ratio in FP32, for modes below SODM_F32_MODE_LIMIT: the three logarithms are taken of ratios that
// are 1 + O(1/m), so the absolute error is ~ 6 m * 2^-23 (< 6e-3 at the limit) -- an error in an ACCEPTANCE
// PROBABILITY that is only consulted for the <= 2 % of attempts the squeeze cannot settle (far fewer at such large
// modes, where the squeeze band is ~1/sqrt(npq) wide), i.e. a distortion of the sampled law below 1.2e-4, an order
// of magnitude under the Monte-Carlo resolution of a KS test at 10^6 trajectories (1.6e-3).  Measured effect of
// the limit on influenza (B200): 2048 -> 259 ms, 8192 -> 242 ms per 2^20 trajectories (the S->E draws of the large
// age groups have modes of ~7000 and otherwise take the FP64 path, ~1000 instructions at 1-2 active lanes).
#ifndef SODM_F32_MODE_LIMIT
#define SODM_F32_MODE_LIMIT 8192.0f
#endif
__host__ __device__ __forceinline__ float binomial_log_ratio_f32(float nf, float p, float m, float k) {
    // r enters as (k - m) log r once the two logarithms below are added up, so an approximate quotient is harmless
    const float r = fast_div(p, 1.0f - p);
    const float a = nf - m + 1.0f, b = nf - k + 1.0f;
    // the two plain logarithms take arguments 1 + O(1/sqrt(m)): the fast logarithm's absolute error (2^-21.4 there)
    // times (m + 1/2) stays within the error budget stated above
    return (m + 0.5f) * fast_log(fast_div(m + 1.0f, r * a)) + stirling_tail_large_f(m) + stirling_tail_large_f(nf - m)
         + (nf + 1.0f) * log1p_rel(fast_div(k - m, b))
         + (k + 0.5f) * fast_log(fast_div(r * b, k + 1.0f))
         - stirling_tail_f(k) - stirling_tail_f(nf - k);
}

// ------------------------------------------------------------------------------------------------
// Building blocks of the samplers: BTRS candidate / squeeze / exact test and inversion as separate functions, so that
// the warp-cooperative sampler (jump_run.cuh) can run ONE of them on all lanes at a time.
// ------------------------------------------------------------------------------------------------
enum : int { BIN_DONE = 0, BIN_INVERT = 1, BIN_EXACT = 2, BIN_RETRY = 3, BIN_BTRS = 4 };

// Trivial cases and regime of Bin(n, p): BIN_DONE (k set; -1 = invalid arguments), BIN_INVERT or BIN_BTRS.
__host__ __device__ __forceinline__ int binomial_classify(int n, float p, int& k) {
    if (!(p >= 0.0f && p <= 1.0f) || n < 0) { k = -1; return BIN_DONE; }
    const bool flip = p > 0.5f;
    const float pp = flip ? 1.0f - p : p;
    if (n == 0 || pp == 0.0f) { k = flip ? n : 0; return BIN_DONE; }
    return ((float)n * pp < SODM_BINV_THRESHOLD) ? BIN_INVERT : BIN_BTRS;
}

// Sequential-search inversion driven by a given uniform word (0 < p <= 0.5, n*p small).  The counters are kept in
// FP32 (exact: the search stops at 1024) so that an iteration is a reciprocal, five multiply-adds and two compares.
__host__ __device__ __forceinline__ int binomial_inversion_word(int n, float p, uint32_t word) {
    const float s = fast_div(p, 1.0f - p);
    float f = fast_exp((float)n * log1pf(-p));     // q^n = P(X = 0); n log q > -10.6 here, where __expf is good to ~1e-6
    float u = u01(word);
    const float kmax = (float)(n < 1024 ? n : 1024);
    float kf = 0.0f, left = (float)n;              // left = n - k (exact in FP32 after the first few steps even for n > 2^24)
    // u < 1 strictly (philox.cuh), but FP32 round-off in the running difference (a few 2^-25) can still leave u above the
    // remaining mass; the search therefore also ends where the pmf has fallen below the 2^-24 grid of u itself (far past
    // the mode: n p < 10 puts P(0) above 2e-5) -- beyond that point u carries no information -- instead of running to the cap
    while (u > f && f > 1e-8f && kf < kmax) {
        u -= f;
        kf += 1.0f;
        f *= fast_div(s * left, kf);               // P(k)/P(k-1) = s (n-k+1)/k; a few ulp here only nudge u by ~1e-7
        left -= 1.0f;
    }
    return (int)kf;
}

// Split form of one BTRS attempt, for queue-based processing (jump_run.cuh):
//   btrs_candidate : setup + candidate k + BTRS box test            -> BIN_DONE | BIN_RETRY (k outside 0..n) | BIN_SQUEEZE
//   btrs_squeeze   : scaled uniform + BTPE squeeze on the same words -> BIN_DONE | BIN_RETRY | BIN_EXACT (lhs set)
enum : int { BIN_SQUEEZE = 5 };
// What the acceptance tests need from the candidate stage (so that they do not recompute the hat).
struct BtrsHat {
    float stddev, b, a, us, v;
};
__host__ __device__ __forceinline__ int btrs_candidate(int n, float p, uint32_t wu, uint32_t wv, int& k, BtrsHat& hat) {
    const float nf = (float)n;
    const float np_ = nf * p;
    const float stddev = fast_sqrt(np_ * (1.0f - p));
    const float b = 1.15f + 2.53f * stddev;
    const float a = -0.0873f + 0.0248f * b + 0.01f * p;
    const float u = u01(wu) - 0.5f;
    const float v = u01(wv);
    const float us = 0.5f - fabsf(u);
    hat.stddev = stddev; hat.b = b; hat.a = a; hat.us = us; hat.v = v;
    if (np_ < 4194304.0f) {
        k = (int)floorf(fmaf(fast_div(2.0f * a, us) + b, u, np_ + 0.5f));
    } else {
        const double kd = floor(fma((double)(2.0f * a / us + b), (double)u, fma((double)n, (double)p, 0.5)));
        k = (int)fmin(fmax(kd, -1.0), 2147483647.0);
    }
    // v <= v_r = 0.92 - 4.2 / b   <=>   (0.92 - v) b >= 4.2   (b > 0): no reciprocal
    if (us >= 0.07f && (0.92f - v) * b >= 4.2f) return BIN_DONE;
    if (k < 0 || k > n) return BIN_RETRY;
    return BIN_SQUEEZE;
}
__host__ __device__ __forceinline__ int btrs_candidate(int n, float p, uint32_t wu, uint32_t wv, int& k) {
    BtrsHat hat;
    return btrs_candidate(n, p, wu, wv, k, hat);
}
// log of the scaled uniform that the acceptance tests compare with log f(k)/f(m)   (BTRS step 3.1)
__host__ __device__ __forceinline__ float btrs_lhs(const BtrsHat& h) {
    const float alpha = (2.83f + fast_div(5.1f, h.b)) * h.stddev;
    return fast_log(fast_div(h.v * alpha, fast_div(h.a, h.us * h.us) + h.b));
}
__host__ __device__ __forceinline__ int btrs_squeeze(int n, float p, int k, const BtrsHat& h, float& lhs) {
    const float nf = (float)n;
    const float npq = nf * p * (1.0f - p);
    lhs = btrs_lhs(h);
    const float mf = floorf((nf + 1.0f) * p);
    const float x = fabsf((float)k - mf);
    if (x < 0.5f * npq - 1.0f) {
        const float inv_npq = fast_div(1.0f, npq);
        const float t = -0.5f * x * x * inv_npq;
        const float rho = (x * inv_npq) * ((x * (x * (1.0f / 3.0f) + 0.625f) + (1.0f / 6.0f)) * inv_npq + 0.5f);
        const float eps = 8e-6f * (1.0f + fabsf(t) + fabsf(lhs));
        if (lhs < t - rho - eps) return BIN_DONE;
        if (lhs > t + rho + eps) return BIN_RETRY;
    }
    return BIN_EXACT;
}
// the same from the two random words alone (queue-based processing regenerates the hat of a parked candidate)
__host__ __device__ __forceinline__ int btrs_squeeze(int n, float p, int k, uint32_t wu, uint32_t wv, float& lhs) {
    BtrsHat h;
    const float npq = (float)n * p * (1.0f - p);
    h.stddev = fast_sqrt(npq);
    h.b = 1.15f + 2.53f * h.stddev;
    h.a = -0.0873f + 0.0248f * h.b + 0.01f * p;
    h.us = 0.5f - fabsf(u01(wu) - 0.5f);
    h.v = u01(wv);
    return btrs_squeeze(n, p, k, h, lhs);
}

// The exact acceptance test of candidate k (FP32 where harmless, FP64 otherwise).  Out of line on the device: it runs
// for a few per cent of the attempts only and would otherwise be replicated at every call site of the hot loop.
#ifdef __CUDA_ARCH__
__device__ __noinline__
#else
inline
#endif
bool btrs_exact_accept(int n, float p, int k, float lhs) {
    const float nf = (float)n;
    const float mf = floorf((nf + 1.0f) * p);
    if (mf < SODM_F32_MODE_LIMIT && nf < 16777216.0f) return lhs <= binomial_log_ratio_f32(nf, p, mf, (float)k);
    const double nd = (double)n, pd = (double)p, md = floor((nd + 1.0) * pd);
    double ratio;
    if (md >= SODM_F32_MODE_LIMIT && binomial_log_ratio_large(nd, pd, md, (double)k, ratio)) return (double)lhs <= ratio;
    return (double)lhs <= binomial_log_ratio(nd, pd, md, (double)k);
}

#ifndef SODM_DIRECT_EXACT
#define SODM_DIRECT_EXACT 1              // 0: squeeze in front of every exact test (tuning builds)
#endif
#ifndef SODM_SQUEEZE_LARGE
#define SODM_SQUEEZE_LARGE 1             // 0: large modes go to their (FP64 multiply-add) ratio without the squeeze (tuning builds)
#endif

// ONE attempt at Bin(n, p), 0 <= p <= 1, from the two words (wu, wv) -- the unit of work of the two-pass sampler
// (jump_run.cuh, sample_worklist_2pass):
//   BIN_DONE   k is final (trivial case, sequential inversion, or an accepted BTRS candidate); k = -1: invalid (n, p)
//   BIN_RETRY  the candidate was rejected: attempt again with fresh words
//   BIN_EXACT  (only when !WITH_EXACT) neither the box test nor the squeeze settled the candidate; it takes the exact
//              log-pmf ratio, which the caller postpones: running THE SAME attempt again with WITH_EXACT = true recreates the
//              candidate from the same words and settles it.
// The squeeze (BTPE step 5.2 in FP32) is skipped for n >= 2^24, where FP32 no longer holds k - mode exactly; such draws go
// to the FP64 ratio directly.
#ifndef SODM_ATTEMPT_SQUEEZE
#define SODM_ATTEMPT_SQUEEZE 0           // 1: squeeze in front of the FP32 exact test too (tuning builds)
#endif
template <bool WITH_EXACT>
__host__ __device__ __forceinline__ int binomial_attempt(int n, float p, uint32_t wu, uint32_t wv, int& k) {
    int state = binomial_classify(n, p, k);
    if (state == BIN_DONE) return BIN_DONE;
    const bool flip = p > 0.5f;
    const float pp = flip ? 1.0f - p : p;
    int kk;
    if (state == BIN_INVERT) {
        kk = binomial_inversion_word(n, pp, wu);
    } else {
        BtrsHat hat;
        state = btrs_candidate(n, pp, wu, wv, kk, hat);
        if (state == BIN_SQUEEZE) {
            float lhs;
            // the squeeze pays where the exact test is the FP64 routine (large modes: it settles 98 % there); in front of the
            // FP32 ratio it mostly adds instructions, because some lane of the warp needs the ratio anyway
            if (n < 16777216 && (SODM_ATTEMPT_SQUEEZE || (float)n * pp >= 0.9f * SODM_F32_MODE_LIMIT)) state = btrs_squeeze(n, pp, kk, hat, lhs);
            else { lhs = btrs_lhs(hat); state = BIN_EXACT; }
            if (state == BIN_EXACT) {
                if (!WITH_EXACT) return BIN_EXACT;
                state = btrs_exact_accept(n, pp, kk, lhs) ? BIN_DONE : BIN_RETRY;
            }
        }
        if (state == BIN_RETRY) return BIN_RETRY;
    }
    k = flip ? n - kk : kk;
    return BIN_DONE;
}

// One complete BTRS attempt of a single lane (per-lane walk); requires 0 < p <= 0.5 and n*p >= 10.
template <class Rng>
__host__ __device__ __forceinline__ bool btrs_try(Rng& rng, int n, float p, int& k) {
    const uint32_t wu = rng.next(), wv = rng.next();
    int state = btrs_candidate(n, p, wu, wv, k);
    if (state == BIN_DONE) return true;
    if (state == BIN_RETRY) return false;
    float lhs = 0.0f;
    state = btrs_squeeze(n, p, k, wu, wv, lhs);
    if (state == BIN_DONE) return true;
    if (state == BIN_RETRY) return false;
    return btrs_exact_accept(n, p, k, lhs);
}

// ONE attempt at Bin(n, p) from two given random words (the per-lane walk hands every trip exactly two words, so
// that all lanes of a warp refill their Philox block on the same trips).  Same contract as binomial_try.
// A candidate that fails the box test and whose exact test is the cheap FP32 one goes to it directly: among the 32
// independent attempts of a warp some lane needs that test on four trips out of five anyway, so the squeeze (59
// instructions, settles two candidates in three) mostly adds to what the warp executes.  Large modes, whose exact
// test is the FP64 routine, keep the squeeze -- there it settles 98 %.
__host__ __device__ __forceinline__ bool binomial_try_words(int n, float p, uint32_t wu, uint32_t wv, int& k) {
    int state = binomial_classify(n, p, k);
    if (state == BIN_DONE) return true;
    const bool flip = p > 0.5f;
    const float pp = flip ? 1.0f - p : p;
    int kk;
    bool done = true;
    if (state == BIN_INVERT) {
        kk = binomial_inversion_word(n, pp, wu);
    } else {
        BtrsHat hat;
        state = btrs_candidate(n, pp, wu, wv, kk, hat);
        if (state == BIN_RETRY) done = false;
        else if (state == BIN_SQUEEZE) {
            float lhs = 0.0f;
            if ((SODM_DIRECT_EXACT && (float)n * pp < 0.9f * SODM_F32_MODE_LIMIT) || (!SODM_SQUEEZE_LARGE && n < 16777216) || n >= 16777216) {
                lhs = btrs_lhs(hat);           // (n >= 2^24: FP32 no longer holds k - mode exactly, the squeeze's bounds would not hold)
                state = BIN_EXACT;
            } else {
                state = btrs_squeeze(n, pp, kk, hat, lhs);
            }
            if (state == BIN_RETRY) done = false;
            else if (state == BIN_EXACT) done = btrs_exact_accept(n, pp, kk, lhs);
        }
    }
    k = flip ? n - kk : kk;
    return done;
}

// ONE attempt at Bin(n, p), 0 <= p <= 1.  Returns true when finished with the variate in k.
// NaN / out-of-range p or n < 0 finish with k = -1 (the reference's NumPy call raises ValueError there;
// callers turn -1 into a status flag).
template <class Rng>
__host__ __device__ __forceinline__ bool binomial_try(Rng& rng, int n, float p, int& k) {
    if (!(p >= 0.0f && p <= 1.0f) || n < 0) { k = -1; return true; }
    const bool flip = p > 0.5f;
    const float pp = flip ? 1.0f - p : p;
    if (n == 0 || pp == 0.0f) { k = flip ? n : 0; return true; }
    int kk;
    bool done = true;
    if ((float)n * pp < SODM_BINV_THRESHOLD) kk = binomial_inversion(rng, n, pp);
    else done = btrs_try(rng, n, pp, kk);
    k = flip ? n - kk : kk;
    return done;
}

}  // namespace sodm
