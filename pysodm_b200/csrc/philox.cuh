// Philox4x32-10 counter-based generator, one independent stream per trajectory.
//
// Replaces the reference's per-draw reseeding of NumPy's global MT19937
// (/root/reference/src/pySODM/models/base.py:481,561,566): there a 32-bit seed per draw is taken
// from the parent's RNG; here the stream of trajectory `g` is the Philox keystream with
//   key     = (seed_lo, seed_hi)
//   counter = (block_index, lane_purpose, g_lo, g_hi)
// so a trajectory's randomness depends only on (seed, global trajectory id) -- results do not change
// with the number of GPUs or the launch geometry.
#pragma once
#include <cstdint>

namespace sodm {

// Rounds of the tau-leap walk's generator.  Philox4x32 is "Crush-resistant" from 7 rounds on (Salmon, Moraes, Dror & Shaw,
// SC'11, table 2: Philox4x32-7 passes the whole of BigCrush); 10 is the published default with a safety margin.  The walk
// draws ~2 x 10^4 words per trajectory and the generator was 12.8 % of its instructions (profiles/r2_tau_leap_ncu_full.txt):
// 7 rounds measure 108.8 against 113.3 ms per 2^20 influenza trajectories (B200, profiles/r2_variants_ab.txt).  Everything else
// (SSA, noise, optimiser moves) keeps 10 rounds; the law is checked by the same chi-square / KS tests either way.
#ifndef SODM_WALK_PHILOX_ROUNDS
#define SODM_WALK_PHILOX_ROUNDS 7
#endif

struct Philox4x32 {
    uint32_t key0, key1;      // seed
    uint32_t ctr0, ctr1;      // running block counter (64 bit)
    uint32_t ctr2, ctr3;      // global trajectory id

    __host__ __device__ __forceinline__ void init(uint64_t seed, uint64_t traj, uint32_t purpose = 0) {
        key0 = (uint32_t)seed;
        key1 = (uint32_t)(seed >> 32);
        ctr0 = 0u;
        ctr1 = purpose;
        ctr2 = (uint32_t)traj;
        ctr3 = (uint32_t)(traj >> 32);
    }

    static __host__ __device__ __forceinline__ void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
        lo = a * b;
        hi = __umulhi(a, b);
#else
        uint64_t p = (uint64_t)a * (uint64_t)b;
        lo = (uint32_t)p;
        hi = (uint32_t)(p >> 32);
#endif
    }

    // Generates the next block of 4 x 32 random bits and advances the counter.
    __host__ __device__ __forceinline__ void next4(uint32_t (&out)[4]) { next4_rounds<10>(out); }
    template <int ROUNDS>
    __host__ __device__ __forceinline__ void next4_rounds(uint32_t (&out)[4]) {
        constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
        constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
        uint32_t c0 = ctr0, c1 = ctr1, c2 = ctr2, c3 = ctr3;
        uint32_t k0 = key0, k1 = key1;
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            uint32_t hi0, lo0, hi1, lo1;
            mulhilo(M0, c0, hi0, lo0);
            mulhilo(M1, c2, hi1, lo1);
            c0 = hi1 ^ c1 ^ k0;
            c1 = lo1;
            c2 = hi0 ^ c3 ^ k1;
            c3 = lo0;
            k0 += W0;
            k1 += W1;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
        if (++ctr0 == 0u) ++ctr1;
    }
};

// Philox2x32-10 (Salmon et al. 2011, the 2 x 32-bit member of the same family): two words from a 64-bit counter and a
// 32-bit key, half the work of a 4x32 block.  The tau-leap sampler calls it once per ATTEMPT with the attempt's
// coordinates (trajectory, step, draw, attempt number) as the counter, so the words of an attempt do not depend on what
// any other lane or any other draw is doing (jump_run.cuh).
template <int ROUNDS>
__host__ __device__ __forceinline__ void philox2x32(uint32_t key, uint32_t c0, uint32_t c1, uint32_t& o0, uint32_t& o1) {
    constexpr uint32_t M = 0xD256D193u, W = 0x9E3779B9u;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        uint32_t hi, lo;
        Philox4x32::mulhilo(M, c0, hi, lo);
        c0 = hi ^ key ^ c1;
        c1 = lo;
        key += W;
    }
    o0 = c0;
    o1 = c1;
}

__host__ __device__ __forceinline__ void philox2x32_10(uint32_t key, uint32_t c0, uint32_t c1, uint32_t& o0, uint32_t& o1) {
    philox2x32<10>(key, c0, c1, o0, o1);
}

// 32-bit key of a launch's tau-leap streams from the 64-bit seed (splitmix64 finaliser, folded)
__host__ __device__ __forceinline__ uint32_t fold_seed(uint64_t seed) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)z ^ (uint32_t)(z >> 32);
}

// 24-bit uniform strictly inside (0, 1): k 2^-24 + 2^-26 for k = 0 .. 2^24 - 1, i.e. [2^-26, 1 - 2^-24] after FP32
// rounding (the offset survives only below 1/2, where it keeps the value off zero).  A half-step offset of 2^-25 would round
// the top value to exactly 1.0f, on which a sequential-search inversion never terminates.
__host__ __device__ __forceinline__ float u01(uint32_t x) {
    return (float)(x >> 8) * 5.9604644775390625e-8f + 1.490116119384765625e-8f;
}

}  // namespace sodm
