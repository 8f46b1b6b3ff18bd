// Ensemble post-processing on the device: observation noise, and sums / order statistics over the trajectories of
// every (row, time) cell -- what the reference's workflows do to a finished ensemble on the host:
//   add_*_noise                         /root/reference/src/pySODM/optimization/utils.py:11-95
//   out[state].mean(dim='draws'), .quantile(dim='draws', q=...)
//                                       /root/reference/tutorials/influenza_1718/calibration.py:231,245-246
// so that an 18 GB ensemble is reduced to a few hundred KB before anything crosses PCIe.
//
// Layout: the simulators' own [cell][trajectory] structure-of-arrays (cell = row * n_time + time, stride `ld`): a
// cell's trajectories are contiguous, so every kernel here streams columns with fully coalesced loads.
//   noise        one thread per element, grid-stride; HBM-bound for Gaussian noise, sampler-bound for Poisson / NB
//   sum          one CTA per cell, 16-byte loads, int64 / FP64 accumulation (exact for integer counts)
//   order stats  one CTA per cell: a min/max pass, then MSB-first radix select with 11-bit digits over the bits in
//                which the cell's values actually differ (typically 2 passes), one shared-memory histogram per
//                distinct prefix among the requested ranks
#include "common.cuh"
#include <type_traits>
#include "noise.cuh"

namespace sodm {

// Poisson / negative-binomial noise: one thread per element, blockIdx.y walks cells, x walks trajectories (consecutive
// threads touch consecutive addresses).  Bound by the samplers (~10^2 instructions per variate), not by HBM.
template <typename Tin, typename Tout, bool NEGBIN>
__global__ void __launch_bounds__(256) count_noise_kernel(const Tin* __restrict__ in, Tout* __restrict__ out, int64_t n_cells,
                                                          int64_t n_traj, int64_t ld, int64_t traj_id0, uint64_t seed,
                                                          double alpha, int32_t* __restrict__ status) {
    int32_t st = 0;
    for (int64_t c = blockIdx.y; c < n_cells; c += gridDim.y) {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_traj; i += (int64_t)gridDim.x * blockDim.x) {
            ElementRng rng;
            rng.init(seed, (uint64_t)(traj_id0 + i), (uint32_t)c);
            out[c * ld + i] = (Tout)count_noise<NEGBIN>(rng, alpha, (double)in[c * ld + i], st);
        }
    }
    if (st && status) atomicOr(status, st);
}

// Gaussian noise: one thread per (trajectory, group of four cells) -- one Philox block and two Box-Muller pairs serve
// four elements, which brings the kernel down to its HBM traffic (read + write of the ensemble).
template <typename Tin, typename Tout, bool RELATIVE>
__global__ void __launch_bounds__(256, 4) gaussian_noise_kernel(const Tin* __restrict__ in, Tout* __restrict__ out, int64_t n_cells,
                                                                int64_t n_traj, int64_t ld, int64_t traj_id0, uint64_t seed,
                                                                double sigma) {
    const int64_t n_groups = (n_cells + 3) / 4;
    for (int64_t g = blockIdx.y; g < n_groups; g += gridDim.y) {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_traj; i += (int64_t)gridDim.x * blockDim.x) {
            Tin x[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) x[q] = 4 * g + q < n_cells ? in[(4 * g + q) * ld + i] : Tin(0);   // four loads in flight
            float z[4];
            normal_quad(seed, (uint64_t)(traj_id0 + i), (uint32_t)g, z);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (4 * g + q >= n_cells) continue;
                if constexpr (std::is_same<Tout, float>::value && sizeof(Tin) == 4) {   // float32 result: no reason to go through FP64
                    const float xf = (float)x[q], sf = (float)sigma;
                    out[(4 * g + q) * ld + i] = (Tout)(xf + (RELATIVE ? fabsf(sf * xf) : sf) * z[q]);
                } else {
                    out[(4 * g + q) * ld + i] = (Tout)gaussian_noise(RELATIVE, sigma, (double)x[q], z[q]);
                }
            }
        }
    }
}

// ---- sums ----------------------------------------------------------------------------------------------------
template <typename T> struct Accum { using type = double; };
template <> struct Accum<int32_t> { using type = long long; };

template <typename T>
__global__ void __launch_bounds__(256) draws_sum_kernel(const T* __restrict__ in, int64_t n_traj, int64_t ld, double* __restrict__ out) {
    using A = typename Accum<T>::type;
    const T* col = in + (int64_t)blockIdx.x * ld;
    A acc = 0;
    constexpr int V = 16 / sizeof(T);
    const bool vec = (((uintptr_t)col) & 15) == 0;
    int64_t i = 0;
    if (vec) {
        const int64_t nv = n_traj / V;
        const uint4* c4 = reinterpret_cast<const uint4*>(col);
        for (int64_t j = threadIdx.x; j < nv; j += blockDim.x) {
            const uint4 q = __ldg(c4 + j);
            const T* e = reinterpret_cast<const T*>(&q);
#pragma unroll
            for (int k = 0; k < V; ++k) acc += (A)e[k];
        }
        i = nv * V;
    }
    for (int64_t j = i + threadIdx.x; j < n_traj; j += blockDim.x) acc += (A)col[j];
    __shared__ A part[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        A s = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += part[w];
        out[blockIdx.x] = (double)s;
    }
}

// ---- order statistics ------------------------------------------------------------------------------------------
template <typename T> struct Radix;
template <> struct Radix<int32_t> {
    using Key = uint32_t;
    static constexpr int BITS = 32;
    __device__ static Key encode(int32_t x) { return (uint32_t)x ^ 0x80000000u; }
    __device__ static double decode(Key k) { return (double)(int32_t)(k ^ 0x80000000u); }
    __device__ static bool is_nan(int32_t) { return false; }
};
template <> struct Radix<float> {
    using Key = uint32_t;
    static constexpr int BITS = 32;
    __device__ static Key encode(float x) { const uint32_t b = __float_as_uint(x); return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u); }
    __device__ static double decode(Key k) { return (double)__uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu)); }
    __device__ static bool is_nan(float x) { return x != x; }
};
template <> struct Radix<double> {
    using Key = unsigned long long;
    static constexpr int BITS = 64;
    __device__ static Key encode(double x) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(x);
        return b ^ ((b >> 63) ? 0xffffffffffffffffull : 0x8000000000000000ull);
    }
    __device__ static double decode(Key k) { return __longlong_as_double((long long)(k ^ ((k >> 63) ? 0x8000000000000000ull : 0xffffffffffffffffull))); }
    __device__ static bool is_nan(double x) { return x != x; }
};

constexpr int SELECT_THREADS = 512;
constexpr int DIGIT_BITS = 11;
constexpr int BINS = 1 << DIGIT_BITS;
constexpr int MAX_RANKS = SODM_MAX_RANKS;

struct RankList {
    int64_t rank[MAX_RANKS];
    int n;
};

// Visit every element of a column: 16-byte loads where the column is aligned, four independent loads in flight per
// thread otherwise.  `f(key)` is called by every thread for each of its elements.
template <typename T, class F>
__device__ __forceinline__ void for_each_key(const T* __restrict__ col, int64_t n, F&& f) {
    using R = Radix<T>;
    constexpr int V = 16 / sizeof(T);
    int64_t done = 0;
    if ((((uintptr_t)col) & 15) == 0) {
        const int64_t nv = n / V;
        const uint4* c4 = reinterpret_cast<const uint4*>(col);
        int64_t j = threadIdx.x;
        for (; j + blockDim.x < nv; j += 2 * blockDim.x) {          // two vectors in flight
            const uint4 q0 = __ldg(c4 + j), q1 = __ldg(c4 + j + blockDim.x);
            const T* e0 = reinterpret_cast<const T*>(&q0);
            const T* e1 = reinterpret_cast<const T*>(&q1);
#pragma unroll
            for (int k = 0; k < V; ++k) f(e0[k]);
#pragma unroll
            for (int k = 0; k < V; ++k) f(e1[k]);
        }
        for (; j < nv; j += blockDim.x) {
            const uint4 q = __ldg(c4 + j);
            const T* e = reinterpret_cast<const T*>(&q);
#pragma unroll
            for (int k = 0; k < V; ++k) f(e[k]);
        }
        done = nv * V;
    }
    for (int64_t j = done + threadIdx.x; j < n; j += blockDim.x) f(col[j]);
    (void)sizeof(R);
}

// `out[j * n_cells + cell]` = the rank[j]-th smallest (0-based) of the cell's n_traj values; NaN for every rank of a
// cell that holds a NaN (np.quantile's behaviour).
//
// One CTA per cell.  Pass 0 finds the smallest and largest key: trajectories of one cell are close together (counts of
// 6.0e6 .. 7.4e6, say), so the keys share their leading bits and the radix search only has to resolve
// bits(max - min) of (key - min) -- two 11-bit passes instead of three for typical epidemic counts, none at all for a
// constant cell -- and the first histogram is spread over its bins instead of hammering one shared-memory address.
// Each further pass histograms one digit of the elements that match a rank's settled prefix; ranks with equal prefixes
// share a histogram.
template <typename T>
__global__ void __launch_bounds__(SELECT_THREADS) draws_select_kernel(const T* __restrict__ in, int64_t n_traj, int64_t ld,
                                                                       const __grid_constant__ RankList ranks, int64_t n_cells,
                                                                       double* __restrict__ out) {
    using R = Radix<T>;
    using Key = typename R::Key;
    __shared__ uint32_t hist[MAX_RANKS][BINS];
    __shared__ Key prefix[MAX_RANKS];            // settled high bits of (key - lo), per GROUP of ranks that share them
    __shared__ Key next_prefix[MAX_RANKS];
    __shared__ int64_t remaining[MAX_RANKS];     // per rank: its rank among the elements that match its group's prefix
    __shared__ int group_of[MAX_RANKS];
    __shared__ int n_groups;
    __shared__ Key warp_lo[SELECT_THREADS / 32], warp_hi[SELECT_THREADS / 32];
    __shared__ int any_nan;
    __shared__ Key lo_s, hi_s;
    const T* col = in + (int64_t)blockIdx.x * ld;
    const int J = ranks.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        n_groups = 1;
        any_nan = 0;
        prefix[0] = 0;
        for (int j = 0; j < J; ++j) { group_of[j] = 0; remaining[j] = ranks.rank[j]; }
    }
    // ---- pass 0: range of the keys ----
    Key lo = ~(Key)0, hi = 0;
    bool saw_nan = false;
    for_each_key(col, n_traj, [&](T x) {
        saw_nan |= R::is_nan(x);
        const Key k = R::encode(x);
        lo = k < lo ? k : lo;
        hi = k > hi ? k : hi;
    });
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const Key l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
    }
    if (lane == 0) { warp_lo[warp] = lo; warp_hi[warp] = hi; }
    if (saw_nan) any_nan = 1;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < SELECT_THREADS / 32; ++w) {
            lo = warp_lo[w] < lo ? warp_lo[w] : lo;
            hi = warp_hi[w] > hi ? warp_hi[w] : hi;
        }
        lo_s = lo;
        hi_s = hi;
    }
    __syncthreads();
    lo = lo_s;
    const Key range = hi_s - lo;
    int total_bits = 0;                          // bits needed for key - lo
    while (total_bits < R::BITS && (range >> total_bits) != 0) ++total_bits;

    int settled = 0;                             // number of high bits (of total_bits) in `prefix`
    while (settled < total_bits) {
        const int bits = (total_bits - settled) < DIGIT_BITS ? (total_bits - settled) : DIGIT_BITS;
        const int shift = total_bits - settled - bits;
        const int G = n_groups;
        for (int b = threadIdx.x; b < G * BINS; b += blockDim.x) (&hist[0][0])[b] = 0;
        __syncthreads();
        const uint32_t mask = (1u << bits) - 1u;
        if (settled == 0 && total_bits <= 5) {
            // at most 32 distinct values (early-epidemic counts, say): every warp would hammer a few shared-memory
            // addresses, so lanes with equal digits elect one of them to add their count
            for_each_key(col, n_traj, [&](T x) {
                const uint32_t digit = (uint32_t)(R::encode(x) - lo);
                const uint32_t same = __match_any_sync(__activemask(), digit);
                if (lane == __ffs(same) - 1) atomicAdd(&hist[0][digit], (uint32_t)__popc(same));
            });
        } else if (settled == 0) {
            for_each_key(col, n_traj, [&](T x) {
                const Key k = R::encode(x) - lo;
                atomicAdd(&hist[0][(uint32_t)(k >> shift) & mask], 1u);
            });
        } else {
            Key p[MAX_RANKS];
#pragma unroll
            for (int g = 0; g < MAX_RANKS; ++g) p[g] = g < G ? prefix[g] : ~(Key)0;
            for_each_key(col, n_traj, [&](T x) {
                const Key k = R::encode(x) - lo;
                const Key high = k >> (shift + bits);
                const uint32_t digit = (uint32_t)(k >> shift) & mask;
#pragma unroll
                for (int g = 0; g < MAX_RANKS; ++g)
                    if (g < G && high == p[g]) atomicAdd(&hist[g][digit], 1u);
            });
        }
        __syncthreads();
        // one warp per rank finds the bin holding it: lane l owns bins [l * BINS/32, (l+1) * BINS/32)
        int64_t my_remaining = 0;
        uint32_t my_digit = 0;
        if (warp < J) {
            const uint32_t* h = hist[group_of[warp]];
            constexpr int PER = BINS / 32;
            uint32_t local = 0;
            for (int b = 0; b < PER; ++b) local += h[lane * PER + b];
            uint32_t incl = local;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            const int64_t want = remaining[warp];
            const int64_t before = (int64_t)incl - local;
            const bool mine = want >= before && want < (int64_t)incl;
            const uint32_t owner = __ballot_sync(0xffffffffu, mine);
            if (mine) {
                int64_t cum = before;
                int b = 0;
                for (; b < PER - 1; ++b) {
                    const uint32_t c = h[lane * PER + b];
                    if (want < cum + c) break;
                    cum += c;
                }
                my_digit = (uint32_t)(lane * PER + b);
                my_remaining = want - cum;
            }
            const int src = owner ? __ffs(owner) - 1 : 0;
            my_digit = __shfl_sync(0xffffffffu, my_digit, src);
            my_remaining = __shfl_sync(0xffffffffu, my_remaining, src);
        }
        __syncthreads();
        if (warp < J && lane == 0) {
            next_prefix[warp] = (settled ? (prefix[group_of[warp]] << bits) : (Key)0) | (Key)my_digit;
            remaining[warp] = my_remaining;
        }
        __syncthreads();
        if (threadIdx.x == 0) {                  // regroup: ranks with equal prefixes share a histogram
            int g_count = 0;
            for (int j = 0; j < J; ++j) {
                int g = -1;
                for (int q = 0; q < g_count; ++q) if (prefix[q] == next_prefix[j]) { g = q; break; }
                if (g < 0) { g = g_count++; prefix[g] = next_prefix[j]; }
                group_of[j] = g;
            }
            n_groups = g_count;
        }
        settled += bits;
        __syncthreads();
    }
    if ((int)threadIdx.x < J) {
        const double v = R::decode(lo + prefix[group_of[threadIdx.x]]);
        out[(int64_t)threadIdx.x * n_cells + blockIdx.x] = any_nan ? __longlong_as_double(0x7ff8000000000000ll) : v;
    }
}

// ---- order statistics of cells whose trajectories are spread over several GPUs ------------------------------------
// The fused kernel above needs a whole column in one place.  When an ensemble is sharded by trajectory, the same radix
// search runs in PHASES whose per-cell tables are combined across ranks by the caller (one NCCL all-reduce each) -- the
// data never moves, only histograms do:
//   minmax    local smallest / largest key and a NaN flag per cell            -> all-reduce MIN / MAX / MAX
//   histogram one digit of the elements matching each rank's settled prefix   -> all-reduce SUM   ([cell][rank][bin])
//   pick      (identical on every rank) the bin holding each requested rank; settles up to 11 more bits
//   decode    keys -> values
// Keys travel as int64 with the sign bit flipped, so that signed MIN / MAX order them like the unsigned keys.
struct SelectState {           // device arrays, all-reduced / updated between phases
    long long* lo;             // [n_cells]            smallest key (biased)
    long long* hi;             // [n_cells]            largest key (biased)
    int32_t* nan_flag;         // [n_cells]
    int32_t* settled;          // [n_cells]            bits of (key - lo) already decided
    long long* prefix;         // [n_cells][n_ranks]   decided high bits
    long long* remaining;      // [n_cells][n_ranks]   rank among the elements that match the prefix
};
constexpr unsigned long long KEY_BIAS = 0x8000000000000000ull;

__device__ __forceinline__ int bits_of(unsigned long long range) {
    int b = 0;
    while (b < 64 && (range >> b) != 0) ++b;
    return b;
}

template <typename T>
__global__ void __launch_bounds__(SELECT_THREADS) draws_minmax_kernel(const T* __restrict__ in, int64_t n_traj, int64_t ld, SelectState st) {
    using R = Radix<T>;
    using Key = typename R::Key;
    __shared__ Key warp_lo[SELECT_THREADS / 32], warp_hi[SELECT_THREADS / 32];
    __shared__ int any_nan;
    if (threadIdx.x == 0) any_nan = 0;
    __syncthreads();
    const T* col = in + (int64_t)blockIdx.x * ld;
    Key lo = ~(Key)0, hi = 0;
    bool saw_nan = false;
    for_each_key(col, n_traj, [&](T x) {
        saw_nan |= R::is_nan(x);
        const Key k = R::encode(x);
        lo = k < lo ? k : lo;
        hi = k > hi ? k : hi;
    });
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const Key l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
    }
    if ((threadIdx.x & 31) == 0) { warp_lo[threadIdx.x >> 5] = lo; warp_hi[threadIdx.x >> 5] = hi; }
    if (saw_nan) any_nan = 1;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < SELECT_THREADS / 32; ++w) {
            lo = warp_lo[w] < lo ? warp_lo[w] : lo;
            hi = warp_hi[w] > hi ? warp_hi[w] : hi;
        }
        // an empty shard contributes the neutral elements of MIN / MAX
        st.lo[blockIdx.x] = (long long)(((unsigned long long)lo) ^ KEY_BIAS);
        st.hi[blockIdx.x] = (long long)(((unsigned long long)hi) ^ KEY_BIAS);
        if (n_traj == 0) { st.lo[blockIdx.x] = 0x7fffffffffffffffll; st.hi[blockIdx.x] = (long long)0x8000000000000000ull; }
        st.nan_flag[blockIdx.x] = any_nan;
    }
}

// hist[(cell * n_ranks + j) * BINS + b]; cells that are already settled are skipped without touching their column
template <typename T>
__global__ void __launch_bounds__(SELECT_THREADS) draws_histogram_kernel(const T* __restrict__ in, int64_t n_traj, int64_t ld,
                                                                          SelectState st, int n_ranks, uint32_t* __restrict__ hist) {
    using R = Radix<T>;
    using Key = typename R::Key;
    __shared__ uint32_t h[MAX_RANKS][BINS];
    const int64_t c = blockIdx.x;
    const unsigned long long lo = ((unsigned long long)st.lo[c]) ^ KEY_BIAS, hi = ((unsigned long long)st.hi[c]) ^ KEY_BIAS;
    const int total_bits = bits_of(hi - lo), settled = st.settled[c];
    uint32_t* out = hist + c * n_ranks * BINS;
    if (settled >= total_bits) return;                                  // uniform over the block
    const int bits = (total_bits - settled) < DIGIT_BITS ? (total_bits - settled) : DIGIT_BITS;
    const int shift = total_bits - settled - bits;
    for (int b = threadIdx.x; b < n_ranks * BINS; b += blockDim.x) (&h[0][0])[b] = 0;
    __syncthreads();
    const uint32_t mask = (1u << bits) - 1u;
    const T* col = in + c * ld;
    if (settled == 0) {                                                 // every rank shares the empty prefix: slot 0
        for_each_key(col, n_traj, [&](T x) {
            const unsigned long long k = (unsigned long long)R::encode(x) - lo;
            atomicAdd(&h[0][(uint32_t)(k >> shift) & mask], 1u);
        });
    } else {
        unsigned long long p[MAX_RANKS];
#pragma unroll
        for (int j = 0; j < MAX_RANKS; ++j) p[j] = j < n_ranks ? (unsigned long long)st.prefix[c * n_ranks + j] : ~0ull;
        for_each_key(col, n_traj, [&](T x) {
            const unsigned long long k = (unsigned long long)R::encode(x) - lo;
            const unsigned long long high = k >> (shift + bits);
            const uint32_t digit = (uint32_t)(k >> shift) & mask;
            bool done = false;                                          // ranks with equal prefixes: count once, in the first slot
#pragma unroll
            for (int j = 0; j < MAX_RANKS; ++j)
                if (j < n_ranks && !done && high == p[j]) { atomicAdd(&h[j][digit], 1u); done = true; }
        });
    }
    __syncthreads();
    for (int b = threadIdx.x; b < n_ranks * BINS; b += blockDim.x) out[b] = (&h[0][0])[b];
    (void)sizeof(Key);
}

// one warp per (cell, rank): the bin of the all-reduced histogram that holds the rank
__global__ void __launch_bounds__(128) draws_pick_kernel(SelectState st, int n_ranks, int64_t n_cells, const uint32_t* __restrict__ hist) {
    const int64_t item = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (item >= n_cells * n_ranks) return;
    const int64_t c = item / n_ranks;
    const int j = (int)(item % n_ranks);
    const unsigned long long lo = ((unsigned long long)st.lo[c]) ^ KEY_BIAS, hi = ((unsigned long long)st.hi[c]) ^ KEY_BIAS;
    const int total_bits = bits_of(hi - lo), settled = st.settled[c];
    if (settled >= total_bits) return;
    const int bits = (total_bits - settled) < DIGIT_BITS ? (total_bits - settled) : DIGIT_BITS;
    // the histogram of this rank's prefix lives in the FIRST slot with that prefix
    int slot = 0;
    if (settled > 0) {
        const long long mine = st.prefix[c * n_ranks + j];
        slot = j;
        for (int q = 0; q < j; ++q) if (st.prefix[c * n_ranks + q] == mine) { slot = q; break; }
    }
    const uint32_t* h = hist + (c * n_ranks + slot) * BINS;
    constexpr int PER = BINS / 32;
    unsigned long long local = 0;
    for (int b = 0; b < PER; ++b) local += h[lane * PER + b];
    unsigned long long incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    const long long want = st.remaining[c * n_ranks + j];
    const long long before = (long long)(incl - local);
    if (want >= before && want < (long long)incl) {
        long long cum = before;
        int b = 0;
        for (; b < PER - 1; ++b) {
            const uint32_t cnt = h[lane * PER + b];
            if (want < cum + cnt) break;
            cum += cnt;
        }
        const unsigned long long old = settled ? (unsigned long long)st.prefix[c * n_ranks + j] : 0ull;
        // written to a shadow slot and published by draws_advance_kernel: other warps of this cell still read `prefix`
        st.prefix[(n_cells + c) * n_ranks + j] = (long long)((old << bits) | (unsigned long long)(lane * PER + b));
        st.remaining[c * n_ranks + j] = want - cum;
    }
}

__global__ void draws_advance_kernel(SelectState st, int n_ranks, int64_t n_cells) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const unsigned long long lo = ((unsigned long long)st.lo[c]) ^ KEY_BIAS, hi = ((unsigned long long)st.hi[c]) ^ KEY_BIAS;
    const int total_bits = bits_of(hi - lo), settled = st.settled[c];
    if (settled >= total_bits) return;
    const int bits = (total_bits - settled) < DIGIT_BITS ? (total_bits - settled) : DIGIT_BITS;
    for (int j = 0; j < n_ranks; ++j) st.prefix[c * n_ranks + j] = st.prefix[(n_cells + c) * n_ranks + j];
    st.settled[c] = settled + bits;
}

template <typename T>
__global__ void draws_decode_kernel(SelectState st, int n_ranks, int64_t n_cells, double* __restrict__ out) {
    using R = Radix<T>;
    const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= n_cells * n_ranks) return;
    const int64_t c = item / n_ranks;
    const int j = (int)(item % n_ranks);
    const unsigned long long lo = ((unsigned long long)st.lo[c]) ^ KEY_BIAS;
    const unsigned long long key = lo + (st.settled[c] ? (unsigned long long)st.prefix[c * n_ranks + j] : 0ull);
    out[(int64_t)j * n_cells + c] = st.nan_flag[c] ? __longlong_as_double(0x7ff8000000000000ll) : R::decode((typename R::Key)key);
}

template <typename Tin, typename Tout>
static int launch_noise_typed(const sodm_noise_args* a, cudaStream_t s) {
    const Tin* in = (const Tin*)a->in;
    Tout* out = (Tout*)a->out;
    const unsigned gx = (unsigned)((a->n_traj + 255) / 256 > 4096 ? 4096 : (a->n_traj + 255) / 256);
    const bool gaussian = a->kind == SODM_NOISE_GAUSSIAN_ABS || a->kind == SODM_NOISE_GAUSSIAN_REL;
    const int64_t rows = gaussian ? (a->n_cells + 3) / 4 : a->n_cells;
    // ~32 blocks per SM in total: a block walks many cells for its 256 trajectories instead of 10^7 one-element blocks
    int64_t gy = (148 * 32 + gx - 1) / gx;
    gy = gy > rows ? rows : gy;
    const dim3 grid(gx, (unsigned)gy);
    switch (a->kind) {
        case SODM_NOISE_POISSON:
            count_noise_kernel<Tin, Tout, false><<<grid, 256, 0, s>>>(in, out, a->n_cells, a->n_traj, a->ld, a->traj_id0, a->seed, a->param, a->status); break;
        case SODM_NOISE_NEGBIN:
            count_noise_kernel<Tin, Tout, true><<<grid, 256, 0, s>>>(in, out, a->n_cells, a->n_traj, a->ld, a->traj_id0, a->seed, a->param, a->status); break;
        case SODM_NOISE_GAUSSIAN_ABS:
            gaussian_noise_kernel<Tin, Tout, false><<<grid, 256, 0, s>>>(in, out, a->n_cells, a->n_traj, a->ld, a->traj_id0, a->seed, a->param); break;
        default:
            gaussian_noise_kernel<Tin, Tout, true><<<grid, 256, 0, s>>>(in, out, a->n_cells, a->n_traj, a->ld, a->traj_id0, a->seed, a->param); break;
    }
    SODM_CUDA_CHECK(cudaGetLastError());
    return SODM_OK;
}

template <typename Tin>
static int launch_noise(const sodm_noise_args* a, cudaStream_t s) {
    switch (a->out_dtype) {
        case SODM_I32: return launch_noise_typed<Tin, int32_t>(a, s);
        case SODM_F32: return launch_noise_typed<Tin, float>(a, s);
        case SODM_F64: return launch_noise_typed<Tin, double>(a, s);
    }
    return fail(SODM_EINVAL, "sodm_noise: unknown out_dtype %d", a->out_dtype);
}

}  // namespace sodm

using namespace sodm;

static int check_columns(const char* who, const void* in, int32_t dtype, int64_t n_cells, int64_t n_traj, int64_t ld) {
    if (n_cells < 0 || n_traj < 0) return fail(SODM_EINVAL, "%s: negative size", who);
    if (dtype != SODM_I32 && dtype != SODM_F32 && dtype != SODM_F64) return fail(SODM_EINVAL, "%s: unknown dtype %d", who, dtype);
    if (n_cells && n_traj && !in) return fail(SODM_EINVAL, "%s: null input", who);
    if (ld < n_traj) return fail(SODM_EINVAL, "%s: ld %lld < n_traj %lld", who, (long long)ld, (long long)n_traj);
    return SODM_OK;
}

extern "C" int sodm_noise(const sodm_noise_args* a) {
    if (!a) return fail(SODM_EINVAL, "sodm_noise: null args");
    int rc = check_columns("sodm_noise", a->in, a->in_dtype, a->n_cells, a->n_traj, a->ld);
    if (rc != SODM_OK) return rc;
    if (a->kind < SODM_NOISE_POISSON || a->kind > SODM_NOISE_NEGBIN) return fail(SODM_EINVAL, "sodm_noise: unknown kind %d", a->kind);
    if ((a->kind == SODM_NOISE_GAUSSIAN_ABS || a->kind == SODM_NOISE_GAUSSIAN_REL) && !(a->param >= 0.0))
        return fail(SODM_EINVAL, "sodm_noise: scale < 0");                         // NumPy's message (utils.py:62-64)
    if (a->kind == SODM_NOISE_NEGBIN && !(a->param > 0.0))
        return fail(SODM_EINVAL, "sodm_noise: alpha must be > 0 (n = 1/alpha, utils.py:91)");
    if (a->n_cells == 0 || a->n_traj == 0) return SODM_OK;
    if (!a->out) return fail(SODM_EINVAL, "sodm_noise: null output");
    cudaStream_t s = (cudaStream_t)a->stream;
    switch (a->in_dtype) {
        case SODM_I32: return launch_noise<int32_t>(a, s);
        case SODM_F32: return launch_noise<float>(a, s);
        default:       return launch_noise<double>(a, s);
    }
}

extern "C" int sodm_draws_sum(const void* in, int32_t dtype, int64_t n_cells, int64_t n_traj, int64_t ld, double* sum, void* stream) {
    int rc = check_columns("sodm_draws_sum", in, dtype, n_cells, n_traj, ld);
    if (rc != SODM_OK) return rc;
    if (n_cells == 0) return SODM_OK;
    if (!sum) return fail(SODM_EINVAL, "sodm_draws_sum: null output");
    if (n_cells > 2147483647LL) return fail(SODM_EINVAL, "sodm_draws_sum: too many cells");
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = (unsigned)n_cells;
    switch (dtype) {
        case SODM_I32: draws_sum_kernel<int32_t><<<grid, 256, 0, s>>>((const int32_t*)in, n_traj, ld, sum); break;
        case SODM_F32: draws_sum_kernel<float><<<grid, 256, 0, s>>>((const float*)in, n_traj, ld, sum); break;
        default:       draws_sum_kernel<double><<<grid, 256, 0, s>>>((const double*)in, n_traj, ld, sum); break;
    }
    SODM_CUDA_CHECK(cudaGetLastError());
    return SODM_OK;
}

extern "C" int sodm_draws_order_stats(const void* in, int32_t dtype, int64_t n_cells, int64_t n_traj, int64_t ld,
                                      const int64_t* ranks, int32_t n_ranks, double* out, void* stream) {
    int rc = check_columns("sodm_draws_order_stats", in, dtype, n_cells, n_traj, ld);
    if (rc != SODM_OK) return rc;
    if (n_ranks < 0 || n_ranks > MAX_RANKS) return fail(SODM_EINVAL, "sodm_draws_order_stats: n_ranks %d outside 0..%d", n_ranks, MAX_RANKS);
    if (n_cells == 0 || n_ranks == 0) return SODM_OK;
    if (n_traj == 0) return fail(SODM_EINVAL, "sodm_draws_order_stats: no trajectories");
    if (!ranks || !out) return fail(SODM_EINVAL, "sodm_draws_order_stats: null ranks / output");
    if (n_cells > 2147483647LL) return fail(SODM_EINVAL, "sodm_draws_order_stats: too many cells");
    RankList rl;
    rl.n = n_ranks;
    for (int j = 0; j < n_ranks; ++j) {
        if (ranks[j] < 0 || ranks[j] >= n_traj) return fail(SODM_EINVAL, "sodm_draws_order_stats: rank %lld outside 0..n_traj-1", (long long)ranks[j]);
        rl.rank[j] = ranks[j];
    }
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = (unsigned)n_cells;
    switch (dtype) {
        case SODM_I32: draws_select_kernel<int32_t><<<grid, SELECT_THREADS, 0, s>>>((const int32_t*)in, n_traj, ld, rl, n_cells, out); break;
        case SODM_F32: draws_select_kernel<float><<<grid, SELECT_THREADS, 0, s>>>((const float*)in, n_traj, ld, rl, n_cells, out); break;
        default:       draws_select_kernel<double><<<grid, SELECT_THREADS, 0, s>>>((const double*)in, n_traj, ld, rl, n_cells, out); break;
    }
    SODM_CUDA_CHECK(cudaGetLastError());
    return SODM_OK;
}

// ---- phases of the sharded order statistics (see SelectState) ----------------------------------------------------
static SelectState select_state(const sodm_select_state* p) {
    SelectState st;
    st.lo = (long long*)p->lo; st.hi = (long long*)p->hi; st.nan_flag = p->nan_flag; st.settled = p->settled;
    st.prefix = (long long*)p->prefix; st.remaining = (long long*)p->remaining;
    return st;
}

extern "C" int sodm_draws_select_phase(int32_t phase, const void* in, int32_t dtype, int64_t n_cells, int64_t n_traj, int64_t ld,
                                       int32_t n_ranks, const sodm_select_state* state, uint32_t* hist, double* out, void* stream) {
    int rc = check_columns("sodm_draws_select_phase", in, dtype, n_cells, n_traj, ld);
    if (rc != SODM_OK) return rc;
    if (!state || !state->lo || !state->hi || !state->nan_flag || !state->settled || !state->prefix || !state->remaining)
        return fail(SODM_EINVAL, "sodm_draws_select_phase: incomplete state");
    if (n_ranks < 1 || n_ranks > MAX_RANKS) return fail(SODM_EINVAL, "sodm_draws_select_phase: n_ranks %d outside 1..%d", n_ranks, MAX_RANKS);
    if (n_cells == 0) return SODM_OK;
    if (n_cells > 2147483647LL) return fail(SODM_EINVAL, "sodm_draws_select_phase: too many cells");
    cudaStream_t s = (cudaStream_t)stream;
    const SelectState st = select_state(state);
    const unsigned grid = (unsigned)n_cells;
    switch (phase) {
        case SODM_SELECT_MINMAX:
            switch (dtype) {
                case SODM_I32: draws_minmax_kernel<int32_t><<<grid, SELECT_THREADS, 0, s>>>((const int32_t*)in, n_traj, ld, st); break;
                case SODM_F32: draws_minmax_kernel<float><<<grid, SELECT_THREADS, 0, s>>>((const float*)in, n_traj, ld, st); break;
                default:       draws_minmax_kernel<double><<<grid, SELECT_THREADS, 0, s>>>((const double*)in, n_traj, ld, st); break;
            }
            break;
        case SODM_SELECT_HISTOGRAM:
            if (!hist) return fail(SODM_EINVAL, "sodm_draws_select_phase: null histogram");
            switch (dtype) {
                case SODM_I32: draws_histogram_kernel<int32_t><<<grid, SELECT_THREADS, 0, s>>>((const int32_t*)in, n_traj, ld, st, n_ranks, hist); break;
                case SODM_F32: draws_histogram_kernel<float><<<grid, SELECT_THREADS, 0, s>>>((const float*)in, n_traj, ld, st, n_ranks, hist); break;
                default:       draws_histogram_kernel<double><<<grid, SELECT_THREADS, 0, s>>>((const double*)in, n_traj, ld, st, n_ranks, hist); break;
            }
            break;
        case SODM_SELECT_PICK: {
            if (!hist) return fail(SODM_EINVAL, "sodm_draws_select_phase: null histogram");
            const int64_t warps = n_cells * n_ranks;
            draws_pick_kernel<<<(unsigned)((warps * 32 + 127) / 128), 128, 0, s>>>(st, n_ranks, n_cells, hist);
            draws_advance_kernel<<<(unsigned)((n_cells + 127) / 128), 128, 0, s>>>(st, n_ranks, n_cells);
            break;
        }
        case SODM_SELECT_DECODE: {
            if (!out) return fail(SODM_EINVAL, "sodm_draws_select_phase: null output");
            const unsigned g = (unsigned)((n_cells * n_ranks + 127) / 128);
            switch (dtype) {
                case SODM_I32: draws_decode_kernel<int32_t><<<g, 128, 0, s>>>(st, n_ranks, n_cells, out); break;
                case SODM_F32: draws_decode_kernel<float><<<g, 128, 0, s>>>(st, n_ranks, n_cells, out); break;
                default:       draws_decode_kernel<double><<<g, 128, 0, s>>>(st, n_ranks, n_cells, out); break;
            }
            break;
        }
        default: return fail(SODM_EINVAL, "sodm_draws_select_phase: unknown phase %d", phase);
    }
    SODM_CUDA_CHECK(cudaGetLastError());
    return SODM_OK;
}
