// Per-trajectory drivers for the jump-process steppers (tau-leap and SSA).
//
// Replaces the body of `JumpProcess._solve_discrete` + `_create_fun.func`
// (/root/reference/src/pySODM/models/base.py:352-440) for one trajectory: the time loop
// `while t < t_eval[-1]: y, tau = fun(t, y); t = t + tau` with `t` accumulated in FP64 (:428-434), and
// the final `np.interp` of every state row onto `t_eval = arange(t0, t1 + 1)` (:435-438, :449), which is
// done here as a STREAMING interpolation between the last two step states, so nothing but the output
// grid is ever stored.  The functions are __host__ __device__ so the very same code can be run on the
// CPU by the test-only harness under tests/hostcheck (statistical checks where no GPU is present).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include "jump_models.cuh"

namespace sodm {

constexpr int MAX_PARAM = 64;   // flat parameter elements passed by value
constexpr int MAX_STATE = 64;   // flat state elements of a jump model

enum : uint32_t { COMPAT_STAY_QUIRK = 1u, COMPAT_SSA_FIRST_MATCH = 2u, TUNE_COOPERATIVE = 256u };

struct JumpLaunch {
    int64_t n_traj;            // trajectories in this launch
    int64_t traj_id0;          // global id of trajectory 0 (Philox stream key; GPU-count invariant)
    int64_t ld_in;             // row stride (elements) of y0 / var_values
    int64_t ld_out;            // row stride (elements) of out
    uint64_t seed;
    const double* y0;          // [NS][ld_in], or [NS] when y0_bcast
    int32_t y0_bcast;
    double par[MAX_PARAM];     // broadcast value of every flat parameter element
    int8_t var_row[MAX_PARAM]; // >= 0: element varies per trajectory, row of var_values
    const double* var_values;  // [n_var][ld_in]
    const double* tdpf_table;  // [n_tdpf_values][TDPF_COUNT] (device: staged as FP32 in shared memory)
    const int32_t* tdpf_index; // [n_steps] value in force during step k
    int32_t n_tdpf_values;
    int32_t n_steps;           // tau-leap: number of steps (host replays the FP64 accumulation of t)
    int32_t steps_per_out;     // aligned tau-leap: an output every this many steps; 0 = interpolate
    int32_t n_out;             // output grid points t0, t0+1, .., t0+n_out-1
    double t0;
    double tau;
    void* out;                 // [n_rows][n_out][ld_out]
    int8_t out_row[MAX_STATE]; // state s -> output row, or -1 (state not selected)
    int64_t out_off[MAX_STATE];// state s -> element offset of its output row (row * n_out * ld_out), or -1
    int32_t* status;           // [n_traj] or null
    uint32_t compat;
    int64_t max_events;        // SSA event budget per trajectory
    int32_t lane_shift;        // tau-leap walk: a warp carries 32 >> lane_shift trajectories (small ensembles, jump_kernels.cuh)
};

template <class M>
__host__ __device__ __forceinline__ void load_trajectory(const JumpLaunch& L, int64_t i, double* par, int* y,
                                                         uint32_t& st) {
#pragma unroll
    for (int e = 0; e < M::NPAR; ++e) {
        par[e] = L.par[e];
        if (L.var_row[e] >= 0) par[e] = L.var_values[(int64_t)L.var_row[e] * L.ld_in + i];
    }
#pragma unroll
    for (int s = 0; s < M::NS; ++s) {
        const double v = L.y0_bcast ? L.y0[s] : L.y0[(int64_t)s * L.ld_in + i];
        y[s] = (int)v;
        if ((double)y[s] != v) st |= ST_TRUNCATED_Y0;
    }
}

template <class M, typename OutT>
__host__ __device__ __forceinline__ void emit_counts(const JumpLaunch& L, int64_t i, int j, const int* y) {
    OutT* out = (OutT*)L.out + i + (int64_t)j * L.ld_out;        // column j of trajectory i
#pragma unroll
    for (int s = 0; s < M::NS; ++s) {
        if (L.out_off[s] >= 0) out[L.out_off[s]] = (OutT)y[s];
    }
}

// value at grid time g of the straight line through (t_old, y_old), (t_new, y_new) -- np.interp's own
// arithmetic, slope*(x - xp[j]) + fp[j], kept free of FMA contraction so it rounds like NumPy
__host__ __device__ __forceinline__ double lerp_np(double g, double t_old, double t_new, double y_old, double y_new) {
#ifdef __CUDA_ARCH__
    const double slope = __ddiv_rn(__dsub_rn(y_new, y_old), __dsub_rn(t_new, t_old));
    return __dadd_rn(__dmul_rn(slope, __dsub_rn(g, t_old)), y_old);
#else
    volatile double slope = (y_new - y_old) / (t_new - t_old);
    volatile double prod = slope * (g - t_old);
    return prod + y_old;
#endif
}

template <class M, typename OutT>
__host__ __device__ __forceinline__ void emit_interp(const JumpLaunch& L, int64_t i, int j, double g,
                                                     double t_old, double t_new, const int* y_old, const int* y_new) {
    OutT* out = (OutT*)L.out + i;
#pragma unroll
    for (int s = 0; s < M::NS; ++s) {
        if (L.out_row[s] >= 0) {
            const double v = (g == t_new) ? (double)y_new[s]
                                          : lerp_np(g, t_old, t_new, (double)y_old[s], (double)y_new[s]);
            out[((int64_t)L.out_row[s] * L.n_out + j) * L.ld_out] = (OutT)v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// tau-leap
// ------------------------------------------------------------------------------------------------

// The binomial draws of one step of the 32 trajectories of a warp: slot d of lane l holds (n, p) and, once
// drawn, the count T in place of n.  On the device the arrays live in shared memory as [slot][thread of block], so
// a warp touching ONE slot of 32 different lanes (own or not) hits 32 different banks.
struct WorkList {
    int* n;             // warp base pointers: element (d, lane) at [d * stride + lane]
    float* p;
    uint16_t* qm;       // warp queue of items ((d << 5) | lane): BTRS regime from the front, inversion regime from the back
    uint16_t* qs;       // candidates outside the BTRS box, waiting for the squeeze; payload = the attempt's two words
    uint32_t* qs_u;
    uint32_t* qs_v;
    uint16_t* qx;       // candidates the squeeze could not settle, waiting for the exact test; payload = k, log(scaled v)
    uint32_t* qx_k;
    float* qx_l;
    uint8_t* ord;       // per-lane walk: the order in which this lane takes its draws, [q * stride + lane] -> slot (or null)
    int stride;
    int lane;
    unsigned lanes;     // device: lanes of the warp that run a trajectory
    __host__ __device__ __forceinline__ void put(int d, int nd, float pd) { n[d * stride + lane] = nd; p[d * stride + lane] = pd; }
    __host__ __device__ __forceinline__ int get(int d) const { return n[d * stride + lane]; }
};

#ifdef __CUDA_ARCH__
constexpr int COOP_LANES = 32;
__device__ __forceinline__ unsigned coop_ballot(bool x) { return __ballot_sync(0xffffffffu, x); }
__device__ __forceinline__ unsigned coop_lt(int lane) { return (1u << lane) - 1u; }
__device__ __forceinline__ int coop_popc(unsigned m) { return __popc(m); }
__device__ __forceinline__ void coop_sync() { __syncwarp(); }
#else
constexpr int COOP_LANES = 1;        // CPU harness: a "warp" of one lane runs the same algorithm
inline unsigned coop_ballot(bool x) { return x ? 1u : 0u; }
inline unsigned coop_lt(int) { return 0u; }
inline int coop_popc(unsigned m) { return (int)(m & 1u); }
inline void coop_sync() {}
#endif
constexpr int COOP_QS_CAP = 96;
constexpr int COOP_QX_CAP = 96;       // >= COOP_QS_CAP: one squeeze drain can never overflow it

// 4-byte words of queue storage per warp: qm, qs (+2 payload words), qx (+2 payload words)
__host__ __device__ constexpr int coop_queue_words(int ndraw) {
    return (ndraw * 32 + COOP_QS_CAP + COOP_QX_CAP) / 2 + 2 * COOP_QS_CAP + 2 * COOP_QX_CAP;
}

__host__ __device__ __forceinline__ float word_as_float(uint32_t w) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(w);
#else
    float f; memcpy(&f, &w, 4); return f;
#endif
}
__host__ __device__ __forceinline__ uint32_t float_as_word(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t w; memcpy(&w, &f, 4); return w;
#endif
}

// dependency level of every slot: 0 = independent, L = conditional on a slot of level L-1 (bit d of DEP_MASK:
// slot d's n is what slot d-1 left behind)
__host__ __device__ constexpr uint32_t level_mask(uint32_t dep_mask, int ndraw, int level) {
    uint32_t out = 0u;
    int run = 0;
    for (int d = 0; d < ndraw; ++d) {
        run = ((dep_mask >> d) & 1u) ? run + 1 : 0;
        if (run == level) out |= 1u << d;
    }
    return out;
}
__host__ __device__ constexpr int max_level(uint32_t dep_mask, int ndraw) {
    int best = 0, run = 0;
    for (int d = 0; d < ndraw; ++d) {
        run = ((dep_mask >> d) & 1u) ? run + 1 : 0;
        best = run > best ? run : best;
    }
    return best;
}

// Random words of one (trajectory, step): counter-based, so ANY lane can regenerate the words of ANY item.
//   attempts 2c and 2c+1 of slot d : Philox(seed; step, 0x20000000 | d << 16 | c, traj)   (words 0,1 and 2,3)
//   inversion of slot d            : the same block c = 0, word 0
// Nothing random is stored: a pending item carries at most its attempt's two words through the squeeze queue.
struct StepKey {
    uint64_t seed;
    uint64_t traj0;      // global id of lane 0's trajectory
    uint32_t step;
};
#ifdef __CUDA_ARCH__
__device__ __noinline__          // one copy for the first pass and the re-draws: keeps the hot loop's code small
#else
inline
#endif
void philox_block(const StepKey& key, int lane, uint32_t tag, uint32_t (&out)[4]) {
    Philox4x32 g;
    g.init(key.seed, key.traj0 + (uint64_t)lane);
    g.ctr0 = key.step;
    g.ctr1 = tag;
    g.next4(out);
}

// WARP-COOPERATIVE sampler (tuning flag SODM_TUNE_COOPERATIVE; NOT the default -- see the measurements below).
// In a divergent SIMT loop the warp pays (trips) x (union of the branches any lane takes), and a per-lane rejection
// sampler takes a different branch on almost every lane: measured on the per-lane walk, 8-9 of 32 lanes are active
// per instruction (profiles/).  Here the 32 x NDRAW draws of a warp-step are pooled instead: every lane classifies
// its own slots and appends them to warp queues by REGIME; the queues are then processed 32 items at a time, so all
// lanes of the warp run the same code on different items:
//   pass 0   own slots: trivial cases, first-attempt random words, queue by regime (ballot + popc compaction)
//   BTRS     first attempts; accepted -> result; rejected -> compacted in place for the next round with fresh
//            counter-based words; undecided by the squeeze -> exact-test queue
//   exact    log-pmf ratio test on queued candidates; failures rejoin the BTRS queue
//   inverse  sequential-search inversion on the small-mean items
// Items conditional on another slot (DEP_MASK) form a second level that runs after the first is final.
// Measured on B200 (influenza, 2^20 trajectories, profiles/r1_sampler_ab.txt): 22 of 32 lanes active and 40 % fewer
// warp instructions than the walk, but the queue traffic through shared memory, the ballots and the extra Philox
// blocks leave it at 4.6 M traj/s against 6.3 M traj/s for the walk at 32 resident warps per SM; it is kept for
// models whose draws are more expensive relative to the bookkeeping.
template <int NDRAW, uint64_t DEP_MASK64>
__host__ __device__ __forceinline__ void sample_worklist_coop(const StepKey& key, WorkList& w, uint32_t& st) {
    static_assert(NDRAW <= 32, "the cooperative sampler takes at most 32 draws per step");
    constexpr uint32_t DEP_MASK = (uint32_t)DEP_MASK64;
    constexpr int LEVELS = max_level(DEP_MASK, NDRAW) + 1;
    static_assert(LEVELS <= 2, "the cooperative sampler handles compartments with at most two exits");
    constexpr int CAP = NDRAW * COOP_LANES;
    const int S = w.stride, lane = w.lane;
    const unsigned lt = coop_lt(lane);
#pragma unroll
    for (int level = 0; level < LEVELS; ++level) {
        const uint32_t in_level = level_mask(DEP_MASK, NDRAW, level);
        int nb = 0, ni = 0, nx = 0;
        // ---- pass 0: own slots -> trivial results, regime queues ----------------------------------------------------
#pragma unroll 1
        for (int d = 0; d < NDRAW; ++d) {
            if (!((in_level >> d) & 1u)) continue;
            int n = w.n[d * S + lane];
            const float p = w.p[d * S + lane];
            if (level > 0) { n -= w.n[(d - 1) * S + lane]; w.n[d * S + lane] = n; }   // what the previous exit left behind
            int k;
            const int state = binomial_classify(n, p, k);
            if (state == BIN_DONE) {
                if (k < 0) { st |= ST_BAD_PROB; k = 0; }
                w.n[d * S + lane] = k;
            }
            const unsigned mB = coop_ballot(state == BIN_BTRS), mI = coop_ballot(state == BIN_INVERT);
            if (state == BIN_BTRS) w.qm[nb + coop_popc(mB & lt)] = (uint16_t)((d << 5) | lane);
            if (state == BIN_INVERT) w.qm[CAP - 1 - (ni + coop_popc(mI & lt))] = (uint16_t)((d << 5) | lane);
            nb += coop_popc(mB);
            ni += coop_popc(mI);
        }
        coop_sync();
        // ---- BTRS rounds ------------------------------------------------------------------------------------------
        // qm[0, count): items to attempt.  Consumed entries are re-used in place for the items that must be re-drawn
        // (keep <= number consumed, always).  Candidates outside the BTRS box wait in qs for the squeeze, candidates the
        // squeeze cannot settle wait in qx for the exact test; both queues are drained when nearly full and at the end
        // of a round, and what they reject rejoins the re-draw list.
        int count = nb, ns = 0;
        uint32_t attempt = 0u;
        while (count > 0) {
            int keep = 0;
            for (int base = 0; base < count; base += COOP_LANES) {
                const bool valid = base + lane < count;
                int state = BIN_DONE;
                uint16_t e = 0;
                uint32_t wu = 0u, wv = 0u;
                if (valid) {
                    e = w.qm[base + lane];
                    const int d = e >> 5, o = e & 31;
                    const int n = w.n[d * S + o];
                    const float p = w.p[d * S + o];
                    const bool flip = p > 0.5f;
                    uint32_t r[4];
                    philox_block(key, o, 0x20000000u | ((uint32_t)d << 16) | (attempt >> 1), r);
                    wu = (attempt & 1u) ? r[2] : r[0];
                    wv = (attempt & 1u) ? r[3] : r[1];
                    int k;
                    state = btrs_candidate(n, flip ? 1.0f - p : p, wu, wv, k);
                    if (state == BIN_DONE) w.n[d * S + o] = flip ? n - k : k;
                }
                coop_sync();                                 // every lane has read its queue entry
                const unsigned mR = coop_ballot(valid && state == BIN_RETRY), mS = coop_ballot(valid && state == BIN_SQUEEZE);
                if (valid && state == BIN_RETRY) w.qm[keep + coop_popc(mR & lt)] = e;
                if (valid && state == BIN_SQUEEZE) {
                    const int q = ns + coop_popc(mS & lt);
                    w.qs[q] = e;
                    w.qs_u[q] = wu;
                    w.qs_v[q] = wv;
                }
                keep += coop_popc(mR);
                ns += coop_popc(mS);
                coop_sync();
                const bool last = base + COOP_LANES >= count;
                if (ns > COOP_QS_CAP - COOP_LANES || last) {
                    // ---- squeeze ----
                    for (int sb = 0; sb < ns; sb += COOP_LANES) {
                        const bool sv = sb + lane < ns;
                        int sstate = BIN_DONE, k = 0;
                        float lhs = 0.0f;
                        uint16_t se = 0;
                        if (sv) {
                            se = w.qs[sb + lane];
                            const int d = se >> 5, o = se & 31;
                            const int n = w.n[d * S + o];
                            const float p = w.p[d * S + o];
                            const bool flip = p > 0.5f;
                            const float pp = flip ? 1.0f - p : p;
                            const uint32_t su = w.qs_u[sb + lane], sw = w.qs_v[sb + lane];
                            btrs_candidate(n, pp, su, sw, k);
                            sstate = btrs_squeeze(n, pp, k, su, sw, lhs);
                            if (sstate == BIN_DONE) w.n[d * S + o] = flip ? n - k : k;
                        }
                        const unsigned mR2 = coop_ballot(sv && sstate == BIN_RETRY), mX = coop_ballot(sv && sstate == BIN_EXACT);
                        if (sv && sstate == BIN_RETRY) w.qm[keep + coop_popc(mR2 & lt)] = se;
                        if (sv && sstate == BIN_EXACT) {
                            const int q = nx + coop_popc(mX & lt);
                            w.qx[q] = se;
                            w.qx_k[q] = (uint32_t)k;
                            w.qx_l[q] = lhs;
                        }
                        keep += coop_popc(mR2);
                        nx += coop_popc(mX);
                        coop_sync();
                    }
                    ns = 0;
                    // ---- exact log-pmf ratio test (qx can hold a full squeeze queue, so it is drained right after it) ----
                    for (int xb = 0; xb < nx; xb += COOP_LANES) {
                        const bool xv = xb + lane < nx;
                        bool fail = false;
                        uint16_t xe = 0;
                        if (xv) {
                            xe = w.qx[xb + lane];
                            const int d = xe >> 5, o = xe & 31;
                            const int n = w.n[d * S + o];
                            const float p = w.p[d * S + o];
                            const bool flip = p > 0.5f;
                            const int k = (int)w.qx_k[xb + lane];
                            if (btrs_exact_accept(n, flip ? 1.0f - p : p, k, w.qx_l[xb + lane])) w.n[d * S + o] = flip ? n - k : k;
                            else fail = true;
                        }
                        const unsigned mF = coop_ballot(fail);
                        if (fail) w.qm[keep + coop_popc(mF & lt)] = xe;
                        keep += coop_popc(mF);
                    }
                    nx = 0;
                    coop_sync();
                }
            }
            count = keep;
            ++attempt;
        }
        // ---- inversion-regime items ----------------------------------------------------------------------------------
        for (int base = 0; base < ni; base += COOP_LANES) {
            if (base + lane < ni) {
                const uint16_t e = w.qm[CAP - 1 - (base + lane)];
                const int d = e >> 5, o = e & 31;
                const int n = w.n[d * S + o];
                const float p = w.p[d * S + o];
                const bool flip = p > 0.5f;
                uint32_t r[4];
                philox_block(key, o, 0x20000000u | ((uint32_t)d << 16), r);
                const int k = binomial_inversion_word(n, flip ? 1.0f - p : p, r[0]);
                w.n[d * S + o] = flip ? n - k : k;
            }
        }
        coop_sync();
    }
}

// Per-lane walk (default): each lane draws its own slots one after the other with a private Philox stream, retrying
// a rejected attempt on its next trip through the loop while its neighbours move on to their next slot -- the warp
// runs for max_lane(sum of attempts) trips instead of sum_slot(max_lane attempts).  One copy of the sampler code.
// Every trip consumes exactly two words of the lane's Philox stream (an inversion uses one and drops the other), so
// all lanes refill their 4-word block on the same trips: the Philox rounds are issued every second trip for the whole
// warp instead of on almost every trip for a few lanes (ncu: Philox was 18.6 % of the instructions at 6 active lanes).
struct WalkRng {
    Philox4x32 gen;
    uint32_t w0, w1, w2, w3;
    bool second;
};
// Steps between re-sorts of a lane's draw order: a functor declares `static constexpr int REORDER_EVERY = K` to have its
// draws re-sorted every K steps (generated models: nobody tuned their slot order); without the member -- the hand-written
// functors, whose slot order is tuned by hand -- the order is the functor's own.  Measured on influenza (ms per 10^6):
// declaration order 137.4; re-sorted every 4 / 16 / 64 / 128 steps 135.0 / 129.0 / 127.9 / 126.6; sorted ONCE at the first
// step 125.6; best hand-tuned static order 125.8.  Frequent re-sorting loses because every lane then follows its own
// order and the lanes of a warp stop sharing slots; one sort from the (shared or similar) initial state keeps them aligned.
template <class M, class = void> struct functor_reorder { static constexpr int value = 0; };
template <class M> struct functor_reorder<M, decltype((void)M::REORDER_EVERY)> { static constexpr int value = M::REORDER_EVERY; };

// Units of the draw list: a slot together with the slots conditional on it (bit d of DEP_MASK: d belongs to the unit of d-1).
template <int NDRAW, uint64_t DEP_MASK>
struct DrawUnits {
    int n_units = 0;
    int first[NDRAW] = {};
    int size[NDRAW] = {};
    constexpr DrawUnits() {
        for (int d = 0; d < NDRAW; ++d) {
            if ((DEP_MASK >> d) & 1ull) ++size[n_units - 1];
            else { first[n_units] = d; size[n_units] = 1; ++n_units; }
        }
    }
};

// Re-sort this lane's draw order by ascending current mean n min(p, 1 - p) of each unit's first slot.  The ORDER changes
// nothing in the law; it decides how well the lanes of a warp share code paths: lanes start a step together and drift
// apart, so cheap small-mean draws (of mixed regime) go first and the large-mean BTRS draws last, where drifted lanes
// still execute the same instructions (measured on influenza, static orders: 137 -> 126 ms per 10^6 trajectories,
// profiles/r1_sampler_ab.txt).  Sorting by the lane's OWN state keeps a trajectory independent of its warp neighbours
// (results do not depend on sharding or chunking).  Branch-free rank sort: NU^2 compares per re-sort.
template <int NDRAW, uint64_t DEP_MASK>
__host__ __device__ __forceinline__ void reorder_draws(WorkList& w) {
    constexpr DrawUnits<NDRAW, DEP_MASK> U{};
    const int S = w.stride, lane = w.lane;
    float key[NDRAW];
#pragma unroll
    for (int u = 0; u < U.n_units; ++u) {
        const float p = w.p[U.first[u] * S + lane];
        const float k = (float)w.n[U.first[u] * S + lane] * fminf(p, 1.0f - p);
        key[u] = k == k ? k : 0.0f;                              // NaN probabilities are flagged by the sampler; keep a total order
    }
#pragma unroll
    for (int u = 0; u < U.n_units; ++u) {
        int pos = 0;
#pragma unroll
        for (int v = 0; v < U.n_units; ++v)
            if (v != u) pos += (key[v] < key[u] || (key[v] == key[u] && v < u)) ? U.size[v] : 0;
#pragma unroll
        for (int e = 0; e < U.size[u]; ++e) w.ord[(pos + e) * S + lane] = (uint8_t)(U.first[u] + e);
    }
}

template <int NDRAW, uint64_t DEP_MASK, bool REORDERED>
__host__ __device__ __forceinline__ void sample_worklist_walk(WalkRng& rng, WorkList& w, uint32_t& st) {
    const int S = w.stride, lane = w.lane;
    int q = 0, chain = 0;                // chain = what the earlier exits of the same compartment element have taken
    // Every trip is one attempt that consumes exactly two Philox words, and every lane starts a step on a fresh block
    // (the half-used one of the previous step is dropped): the lanes of a warp then refill on the same trips, so the
    // generator -- a sixth of the kernel's instructions -- always runs with the whole warp.  Letting each lane carry
    // its leftover pair across steps halves the active lanes per refill and costs 11 % (profiles/r1_sampler_ab.txt).
    rng.second = false;
    while (q < NDRAW) {
        const int d = REORDERED ? (int)w.ord[q * S + lane] : q;
        int n = w.n[d * S + lane];
        const float p = w.p[d * S + lane];
        const bool dep = (DEP_MASK >> d) & 1ull;
        if (dep) n -= chain;
        if (!rng.second) {
            uint32_t b[4];
            rng.gen.next4(b);
            rng.w0 = b[0]; rng.w1 = b[1]; rng.w2 = b[2]; rng.w3 = b[3];
        }
        const uint32_t wu = rng.second ? rng.w2 : rng.w0, wv = rng.second ? rng.w3 : rng.w1;
        rng.second = !rng.second;
        int k;
        if (binomial_try_words(n, p, wu, wv, k)) {
            if (k < 0) { st |= ST_BAD_PROB; k = 0; }
            w.n[d * S + lane] = k;
            chain = dep ? chain + k : k;
            ++q;
        }
    }
}

// A functor whose rates depend on the time itself (generated models with seasonal forcing, say) declares
// `static constexpr bool USES_TIME = true` and takes the step's start time as a last argument of `propose`.
template <class M, class = void> struct functor_uses_time { static constexpr bool value = false; };
template <class M> struct functor_uses_time<M, decltype((void)M::USES_TIME)> { static constexpr bool value = M::USES_TIME; };

// One trajectory per lane.  `active` = the lane owns a trajectory; the remaining lanes of a partially filled warp
// still take part in the warp-cooperative sampling (they own no items) but neither load nor store.
template <class M, typename OutT, bool ALIGNED, bool COOP>
__host__ __device__ __forceinline__ void tau_leap_trajectory(const JumpLaunch& L, int64_t i, bool active,
                                                             const float* table, WorkList w) {
    double par[M::NPAR];
    int y[M::NS];
    uint32_t st = ST_OK;
    if (active) load_trajectory<M>(L, i, par, y, st);
    else {
#pragma unroll
        for (int e = 0; e < M::NPAR; ++e) par[e] = L.par[e];
#pragma unroll
        for (int s = 0; s < M::NS; ++s) y[s] = 0;
    }
    const bool quirk = (L.compat & COMPAT_STAY_QUIRK) != 0u;
    const float tau = (float)L.tau;
    const typename M::Pre pre = M::prepare(par, tau, quirk);
    WalkRng rng;                           // private stream of the per-lane walk
    rng.gen.init(L.seed, (uint64_t)(L.traj_id0 + i));
    rng.w0 = rng.w1 = rng.w2 = rng.w3 = 0u;
    rng.second = false;
    StepKey key;
    key.seed = L.seed;
    key.traj0 = (uint64_t)(L.traj_id0 + i - w.lane);
    key.step = 0u;

    if (active) emit_counts<M, OutT>(L, i, 0, y);
    int j = 1;
    int y_old[ALIGNED ? 1 : M::NS];
    int until_out = L.steps_per_out;
    double t = L.t0;
    for (int k = 0; k < L.n_steps; ++k) {
        if (!ALIGNED) {
#pragma unroll
            for (int s = 0; s < M::NS; ++s) y_old[s] = y[s];
        }
        const float* tab = M::TDPF_COUNT ? table + (int64_t)L.tdpf_index[k] * M::TDPF_COUNT : table;
        if constexpr (functor_uses_time<M>::value) M::propose(y, pre, tab, tau, quirk, w, ALIGNED ? L.t0 + (double)k * L.tau : t);
        else M::propose(y, pre, tab, tau, quirk, w);
        key.step = (uint32_t)k;
        if constexpr (COOP) sample_worklist_coop<M::NDRAW, (uint64_t)M::DEP_MASK>(key, w, st);
        else {
            constexpr int EVERY = functor_reorder<M>::value;
            if constexpr (EVERY > 0) {
                if (k % EVERY == 0) reorder_draws<M::NDRAW, (uint64_t)M::DEP_MASK>(w);
            }
            sample_worklist_walk<M::NDRAW, (uint64_t)M::DEP_MASK, (EVERY > 0)>(rng, w, st);
        }
        M::apply(y, w);
        if (ALIGNED) {
            if (--until_out == 0) {
                if (active) emit_counts<M, OutT>(L, i, j, y);
                ++j;
                until_out = L.steps_per_out;
            }
        } else {
            // general tau: interpolate onto the integer grid between consecutive step states
            const double t_new = t + L.tau;
            while (j < L.n_out && L.t0 + (double)j <= t_new) {
                if (active) emit_interp<M, OutT>(L, i, j, L.t0 + (double)j, t, t_new, y_old, y);
                ++j;
            }
            t = t_new;
        }
    }
    if (active && L.status) L.status[i] = (int32_t)st;
}

// ------------------------------------------------------------------------------------------------
// Gillespie SSA (base.py:206-281)
// ------------------------------------------------------------------------------------------------
template <class M, typename OutT>
__host__ __device__ __forceinline__ void ssa_trajectory(const JumpLaunch& L, int64_t i) {
    double par[M::NPAR];
    int y[M::NS], y_old[M::NS];
    uint32_t st = ST_OK;
    load_trajectory<M>(L, i, par, y, st);
    const bool first_match = (L.compat & COMPAT_SSA_FIRST_MATCH) != 0u;
    PhiloxWords rng;
    rng.init(L.seed, (uint64_t)(L.traj_id0 + i), 1u);

    emit_counts<M, OutT>(L, i, 0, y);
    const double t_end = L.t0 + (double)(L.n_out - 1);
    double t = L.t0;
    int j = 1;
    int64_t events = 0;
    while (t < t_end) {
        if (events++ >= L.max_events) { st |= ST_MAX_STEPS; break; }
        double a[M::NCH];
        M::propensities(y, par, a);
        double R = 0.0;
#pragma unroll
        for (int c = 0; c < M::NCH; ++c) R += a[c];
#pragma unroll
        for (int s = 0; s < M::NS; ++s) y_old[s] = y[s];
        double step;
        if (R != 0.0) {
            if (!(R > 0.0) || R > 1.0e300) { st |= ST_NONFINITE; break; }
            uint32_t b[4];
            b[0] = rng.next(); b[1] = rng.next(); b[2] = rng.next(); b[3] = rng.next();
            // two 53-bit uniforms in (0,1)
            const double u1 = ((double)(((uint64_t)b[0] << 21) ^ (uint64_t)(b[1] >> 11)) + 0.5) * 1.1102230246251565e-16;
            const double u2 = ((double)(((uint64_t)b[2] << 21) ^ (uint64_t)(b[3] >> 11)) + 0.5) * 1.1102230246251565e-16;
            step = 1.0 / R * log(1.0 / u1);
            // weighted choice: first channel whose cumulative propensity exceeds u2 * R
            const double x = u2 * R;
            double cum = 0.0;
            int ch = -1;
#pragma unroll
            for (int c = 0; c < M::NCH; ++c) {
                cum += a[c];
                if (ch < 0 && x < cum) ch = c;
            }
            if (ch < 0) {                                   // round-off at the top end: last non-zero channel
#pragma unroll
                for (int c = 0; c < M::NCH; ++c) if (a[c] > 0.0) ch = c;
            }
            if (first_match) {                              // reference: list.index(value) (F2)
                double chosen = 0.0;
#pragma unroll
                for (int c = 0; c < M::NCH; ++c) if (c == ch) chosen = a[c];
                int first = -1;
#pragma unroll
                for (int c = 0; c < M::NCH; ++c) if (first < 0 && a[c] == chosen) first = c;
                ch = first;
            }
            M::fire(y, ch);
        } else {
            step = 1.0;                                     // base.py:270-272
        }
        const double t_new = t + step;
        while (j < L.n_out && L.t0 + (double)j <= t_new) {
            emit_interp<M, OutT>(L, i, j, L.t0 + (double)j, t, t_new, y_old, y);
            ++j;
        }
        t = t_new;
    }
    while (j < L.n_out) emit_counts<M, OutT>(L, i, j++, y);   // only after an aborted run: hold the last state
    if (L.status) L.status[i] = (int32_t)st;
}

}  // namespace sodm
