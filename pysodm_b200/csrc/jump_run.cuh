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
#include <type_traits>
#include "jump_models.cuh"

namespace sodm {

constexpr int MAX_PARAM = 64;   // flat parameter elements passed by value
constexpr int MAX_STATE = 64;   // flat state elements of a jump model

enum : uint32_t { COMPAT_STAY_QUIRK = 1u, COMPAT_SSA_FIRST_MATCH = 2u, TUNE_LEGACY_WALK = 256u };

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
        if (y[s] < 0) { st |= ST_BAD_PROB; y[s] = 0; }       // NumPy's binomial raises for n < 0 (base.py:303)
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
// a warp touching ONE slot of its 32 lanes hits 32 different banks.
struct WorkList {
    int* n;             // warp base pointers: element (d, lane) at [d * stride + lane]
    float* p;
    uint8_t* ord;       // the order in which this lane takes its draws, [q * stride + lane] -> slot
    int stride;
    int lane;
    unsigned lanes;     // device: lanes of the warp that run a trajectory
    __host__ __device__ __forceinline__ void put(int d, int nd, float pd) { n[d * stride + lane] = nd; p[d * stride + lane] = pd; }
    __host__ __device__ __forceinline__ int get(int d) const { return n[d * stride + lane]; }
};

#ifdef __CUDA_ARCH__
__device__ __forceinline__ unsigned coop_ballot(bool x) { return __ballot_sync(0xffffffffu, x); }
__device__ __forceinline__ int coop_popc(unsigned m) { return __popc(m); }
__device__ __forceinline__ int low_bit(uint32_t m) { return __ffs((int)m) - 1; }
__device__ __forceinline__ int low_bit(uint64_t m) { return __ffsll((long long)m) - 1; }
#else
inline unsigned coop_ballot(bool x) { return x ? 1u : 0u; }       // CPU harness: a "warp" of one lane runs the same algorithm
inline int coop_popc(unsigned m) { return (int)(m & 1u); }
inline int low_bit(uint32_t m) { return __builtin_ctz(m); }
inline int low_bit(uint64_t m) { return __builtin_ctzll(m); }
#endif

// Draw order of a lane (position -> slot), set once from the lane's OWN first-step means (a trajectory's result must not
// depend on its warp neighbours): independent slots by ascending mean n min(p, 1 - p), then the slots that are conditional
// on another one (bit d of DEP_MASK: slot d's n is what slot d-1 left behind), so that by the time a lane reaches them
// their predecessors are settled.  The ORDER changes nothing in the law; it decides how well the lanes of a warp share
// code paths: trajectories of one ensemble resemble each other, so with a common ordering rule the small-mean draws
// (sequential inversion) of all lanes meet in the first trips of a step and the large-mean draws (BTRS) in the rest,
// whichever lane has fallen behind by a few attempts (measured in round 1 on the per-lane walk: declaration order 137 ms,
// sorted once 126 ms per 10^6 influenza trajectories; re-sorting often is worse because the lanes' orders drift apart).
// A functor may declare `static constexpr int REORDER_EVERY = K` to re-sort every K steps.  Rank sort, NDRAW^2 compares.
template <class M, class = void> struct functor_reorder { static constexpr int value = 0; };
template <class M> struct functor_reorder<M, decltype((void)M::REORDER_EVERY)> { static constexpr int value = M::REORDER_EVERY; };

template <int NDRAW, uint64_t DEP_MASK>
__host__ __device__ __forceinline__ void reorder_draws(WorkList& w) {
    const int S = w.stride, lane = w.lane;
    float key[NDRAW];
#pragma unroll
    for (int d = 0; d < NDRAW; ++d) {
        const float p = w.p[d * S + lane];
        const float k = (float)w.n[d * S + lane] * fminf(p, 1.0f - p);
        key[d] = ((DEP_MASK >> d) & 1ull) ? 3.0e38f : (k == k ? k : 0.0f);     // NaN probabilities are flagged by the sampler; keep a total order
    }
#pragma unroll
    for (int d = 0; d < NDRAW; ++d) {
        int pos = 0;
#pragma unroll
        for (int e = 0; e < NDRAW; ++e)
            if (e != d) pos += (key[e] < key[d] || (key[e] == key[d] && e < d)) ? 1 : 0;
        w.ord[pos * S + lane] = (uint8_t)d;
    }
}

// what the earlier exits of slot d's compartment element have taken (d is conditional: bit d of DEP_MASK)
template <uint64_t DEP_MASK>
__host__ __device__ __forceinline__ int chain_taken(const WorkList& w, int d) {
    int taken = 0, e = d;
    do {
        --e;
        taken += w.n[e * w.stride + w.lane];
    } while ((DEP_MASK >> e) & 1ull);
    return taken;
}

// ------------------------------------------------------------------------------------------------------------------
// DEFERRED-TEST WALK (default sampler).
//
// In a divergent SIMT loop the warp pays (trips) x (union of the branches any lane takes).  The per-lane walk of round 1
// (kept below as the legacy schedule) ran the exact acceptance test of a BTRS candidate right where the box test failed:
// one attempt in three needs it, so SOME lane of 32 needed it on almost every trip and the warp executed the ~130
// instructions of the test at 8 of 32 lanes, next to inversion at 6 lanes -- 17.4 of 32 lanes per instruction overall
// (profiles/r1_mufu_tau_leap_ncu_full.txt).  Here a trip has ONE body for the whole warp:
//   C trip  every lane that still has a draw to do takes the next one in its order and makes ONE attempt from ONE random
//           word: trivial cases and small means (sequential inversion) finish at once; a BTRS candidate that passes the box
//           test (decided from the 8 low bits of the same word) is final; any other candidate is PARKED in registers (its k,
//           its word, its position; up to DEFER_SLOTS of them) and the lane moves on to its next draw -- the draws of a step
//           are independent of each other except for the few conditional ones, which sit last in the order.
//   X trip  every lane that has parked candidates settles the newest one: second word -> full-resolution v -> exact
//           acceptance test (btrs_settle).  Accepted: the count is final.  Rejected: the draw's position returns to the
//           lane's to-do mask and is attempted again on a later C trip.
// The warp votes (ballots) which of the two bodies the next trip runs: X when no lane can make progress otherwise, when
// some lane's parking slots are full, or when at least DEFER_VOTE lanes have something parked.  Every trip consumes exactly
// one word of every lane's Philox stream, used or not, so all lanes refill their 4-word block on the same trips.
// Rejection sampling stays exact under any such schedule: each attempt uses words nobody has looked at before, and whether
// an attempt is accepted depends on its own words only.
// ------------------------------------------------------------------------------------------------------------------
#ifndef SODM_DEFER_SLOTS
#define SODM_DEFER_SLOTS 3
#endif
#ifndef SODM_DEFER_VOTE
#define SODM_DEFER_VOTE 24
#endif
constexpr int DEFER_SLOTS = SODM_DEFER_SLOTS;
static_assert(DEFER_SLOTS >= 1 && DEFER_SLOTS <= 4, "1 to 4 parked candidates per lane");

// One word per trip for every lane; the buffer position is the same on all lanes of a warp by construction.
struct TripRng {
    Philox4x32 gen;
    uint32_t w0, w1, w2, w3;
    int left;
    __host__ __device__ __forceinline__ void init(uint64_t seed, uint64_t traj) {
        gen.init(seed, traj);
        w0 = w1 = w2 = w3 = 0u;
        left = 0;
    }
    __host__ __device__ __forceinline__ uint32_t next() {
        if (left == 0) {
            uint32_t b[4];
            gen.next4(b);
            w0 = b[0]; w1 = b[1]; w2 = b[2]; w3 = b[3];
            left = 4;
        }
        const uint32_t v = w0;
        w0 = w1; w1 = w2; w2 = w3;
        --left;
        return v;
    }
};

template <int NDRAW> struct draw_mask { using type = typename std::conditional<(NDRAW <= 32), uint32_t, uint64_t>::type; };

template <int NDRAW, uint64_t DEP_MASK>
__host__ __device__ __forceinline__ void sample_worklist_defer(TripRng& rng, WorkList& w, uint32_t& st) {
    using Mask = typename draw_mask<NDRAW>::type;
    constexpr Mask ONE = 1;
    constexpr Mask ALL = NDRAW == 8 * (int)sizeof(Mask) ? ~Mask(0) : (ONE << (NDRAW % (8 * (int)sizeof(Mask)))) - ONE;
    const int S = w.stride, lane = w.lane;
    // a conditional draw carries ~n (negative) until its first attempt turns it into n - (what the earlier exits took)
#pragma unroll
    for (int d = 0; d < NDRAW; ++d)
        if ((DEP_MASK >> d) & 1ull) w.n[d * S + lane] = ~w.n[d * S + lane];
    Mask todo = ALL;            // POSITIONS (in the lane's order) still to be attempted
    Mask open = ALL;            // SLOTS whose count is not final yet
    int npend = 0;
    int pk0 = 0, pk1 = 0, pk2 = 0, pk3 = 0;          // parked candidates, newest first: k, word, position (one byte each)
    uint32_t pw0 = 0u, pw1 = 0u, pw2 = 0u, pw3 = 0u, ppos = 0u;
    bool stalled = false;       // the lane's next draw waits for a parked predecessor
    for (;;) {
        const bool can = todo != 0 && npend < DEFER_SLOTS && !stalled;
        const unsigned mc = coop_ballot(can), mp = coop_ballot(npend > 0);
        if ((mc | mp) == 0u) break;
        const uint32_t word = rng.next();
        const bool settle = mp != 0u && (mc == 0u || coop_ballot(npend >= DEFER_SLOTS) != 0u || coop_popc(mp) >= SODM_DEFER_VOTE);
        if (settle) {
            // ---- X trip ------------------------------------------------------------------------------------------
            if (npend > 0) {
                const int pos = (int)(ppos & 255u);
                const int d = (int)w.ord[pos * S + lane];
                const int n = w.n[d * S + lane];
                const float p = w.p[d * S + lane];
                const bool flip = p > 0.5f;
                if (btrs_settle(n, flip ? 1.0f - p : p, pk0, pw0, word)) {
                    w.n[d * S + lane] = flip ? n - pk0 : pk0;
                    open &= ~(ONE << d);
                } else {
                    todo |= ONE << pos;
                }
                pk0 = pk1; pk1 = pk2; pk2 = pk3;
                pw0 = pw1; pw1 = pw2; pw2 = pw3;
                ppos >>= 8;
                --npend;
            }
            stalled = false;
        } else if (can) {
            // ---- C trip ------------------------------------------------------------------------------------------
            const int pos = low_bit(todo);
            const int d = (int)w.ord[pos * S + lane];
            int n = w.n[d * S + lane];
            const float p = w.p[d * S + lane];
            bool go = true;
            if (((DEP_MASK >> d) & 1ull) && n < 0) {                // first attempt of a conditional draw
                if ((open >> (d - 1)) & ONE) { stalled = true; go = false; }     // predecessor parked: wait for an X trip
                else { n = ~n - chain_taken<DEP_MASK>(w, d); w.n[d * S + lane] = n; }
            }
            if (go) {
                todo &= ~(ONE << pos);
                int k;
                int state = binomial_classify(n, p, k);
                const bool flip = p > 0.5f;
                const float pp = flip ? 1.0f - p : p;
                if (state == BIN_DONE) {
                    if (k < 0) { st |= ST_BAD_PROB; k = 0; }
                } else {
                    if (state == BIN_INVERT) { k = binomial_inversion_word(n, pp, word); state = BIN_DONE; }
                    else state = btrs_candidate_word(n, pp, word, k);
                    if (state == BIN_DONE) k = flip ? n - k : k;
                }
                if (state == BIN_DONE) {
                    w.n[d * S + lane] = k;
                    open &= ~(ONE << d);
                } else {                                          // park the candidate
                    pk3 = pk2; pk2 = pk1; pk1 = pk0; pk0 = k;
                    pw3 = pw2; pw2 = pw1; pw1 = pw0; pw0 = word;
                    ppos = (ppos << 8) | (uint32_t)pos;
                    ++npend;
                }
            }
        }
    }
}

// LEGACY per-lane walk (SODM_TUNE_LEGACY_WALK; the default sampler of round 1, kept for A/B measurements and as a second
// schedule of the same law in the tests): each lane draws its slots one after the other with a private Philox stream,
// retrying a rejected attempt on its next trip through the loop while its neighbours move on to their next slot; the
// exact test runs inside the trip that needs it.  Every trip consumes exactly two words (an inversion uses one and drops
// the other), and every lane starts a step on a fresh block, so all lanes refill on the same trips.
struct WalkRng {
    Philox4x32 gen;
    uint32_t w0, w1, w2, w3;
    bool second;
};

template <int NDRAW, uint64_t DEP_MASK>
__host__ __device__ __forceinline__ void sample_worklist_walk(WalkRng& rng, WorkList& w, uint32_t& st) {
    const int S = w.stride, lane = w.lane;
    int q = 0;
    rng.second = false;
    while (q < NDRAW) {
        const int d = (int)w.ord[q * S + lane];
        int n = w.n[d * S + lane];
        const float p = w.p[d * S + lane];
        if ((DEP_MASK >> d) & 1ull) n -= chain_taken<DEP_MASK>(w, d);      // predecessors come earlier in the order: final
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
            ++q;
        }
    }
}

// A functor whose rates depend on the time itself (generated models with seasonal forcing, say) declares
// `static constexpr bool USES_TIME = true` and takes the step's start time as a last argument of `propose`.
template <class M, class = void> struct functor_uses_time { static constexpr bool value = false; };
template <class M> struct functor_uses_time<M, decltype((void)M::USES_TIME)> { static constexpr bool value = M::USES_TIME; };

// One trajectory per lane.  `active` = the lane owns a trajectory; the remaining lanes of a partially filled warp
// still take part in the warp's votes (they have no draws to make) but neither load nor store.
template <class M, typename OutT, bool ALIGNED, bool LEGACY>
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
    TripRng rng;                           // private stream of the trajectory (deferred-test walk)
    rng.init(L.seed, (uint64_t)(L.traj_id0 + i));
    WalkRng wrng;                          // the same stream, consumed in pairs (legacy walk)
    wrng.gen.init(L.seed, (uint64_t)(L.traj_id0 + i));
    wrng.w0 = wrng.w1 = wrng.w2 = wrng.w3 = 0u;
    wrng.second = false;

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
        constexpr int EVERY = functor_reorder<M>::value;
        if (EVERY > 0 ? (k % (EVERY > 0 ? EVERY : 1) == 0) : (k == 0)) reorder_draws<M::NDRAW, (uint64_t)M::DEP_MASK>(w);
        if constexpr (LEGACY) sample_worklist_walk<M::NDRAW, (uint64_t)M::DEP_MASK>(wrng, w, st);
        else sample_worklist_defer<M::NDRAW, (uint64_t)M::DEP_MASK>(rng, w, st);
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
