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

enum : uint32_t { COMPAT_STAY_QUIRK = 1u, COMPAT_SSA_FIRST_MATCH = 2u, TUNE_QUEUE = 256u, TUNE_LATENCY = 512u };

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
    int32_t n_peers;           // further destinations of every output element (peer GPUs' buffers, same layout as `out`)
    void* peer_out[7];
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
    // fused gather: the same coalesced rows go straight into the peers' buffers over NVLink (posted stores; uniform branch)
    for (int p = 0; p < L.n_peers; ++p) {
        OutT* peer = (OutT*)L.peer_out[p] + i + (int64_t)j * L.ld_out;
#pragma unroll
        for (int s = 0; s < M::NS; ++s) {
            if (L.out_off[s] >= 0) peer[L.out_off[s]] = (OutT)y[s];
        }
    }
}

// value at grid time g of the straight line through (t_old, y_old), (t_new, y_new) -- np.interp's own
// arithmetic, slope*(x - xp[j]) + fp[j], kept free of FMA contraction so it rounds like NumPy
__host__ __device__ __forceinline__ double lerp_np(double g, double t_old, double t_new, double y_old, double y_new) {
    // A state that did not change between the two step states (all but one or two of them after an SSA event) has slope
    // 0 / dt -- a quotient that takes the out-of-line slow path of the FP64 division (fdiv.cuh).  Such a state is divided as
    // 1 / dt instead and its value y_old returned as it is: 0 * (g - t_old) + y_old, what np.interp computes, bit for bit.
#ifdef __CUDA_ARCH__
    const double dy = __dsub_rn(y_new, y_old);
    const bool flat = dy == 0.0;
    const double slope = __ddiv_rn(flat ? 1.0 : dy, __dsub_rn(t_new, t_old));
    const double v = __dadd_rn(__dmul_rn(slope, __dsub_rn(g, t_old)), y_old);
    return flat ? y_old : v;
#else
    if (y_new == y_old) return y_old;
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
            for (int p = 0; p < L.n_peers; ++p)
                ((OutT*)L.peer_out[p] + i)[((int64_t)L.out_row[s] * L.n_out + j) * L.ld_out] = (OutT)v;
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
    uint8_t* ord;       // walk: the order in which this lane takes its draws, [q * stride + lane] -> slot
    uint16_t* queue;    // queue sampler: the warp's item array, 2 * NDRAW * 32 entries
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

// PER-LANE WALK: draw order.  A lane takes its draws in an order (position -> slot, `ord`) set once from its OWN first-step
// means (a trajectory's result must not depend on its warp neighbours).  The ORDER changes nothing in the law; it decides how
// well the lanes of a warp share code paths: trajectories of one ensemble resemble each other, so with a common ordering
// rule the small-mean draws (sequential inversion) of all lanes meet in the first trips of a step and the large-mean draws
// (BTRS) in the rest, whichever lane has fallen behind by a few attempts (round 1, influenza: declaration order 137 ms,
// sorted once at the first step 125.6 ms, re-sorted every 16 steps 129 ms per 10^6 trajectories -- the lanes' orders drift
// apart when re-sorted often).  Every model uses this generic rule; round 1's hand-tuned table for influenza (125.8 ms) was
// fitted to the benchmark's own trajectory and is gone.
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
// QUEUE SAMPLER (SODM_TUNE_QUEUE; same law as the walk, different random streams).
//
// In a divergent SIMT loop the warp pays (trips) x (union of the branches any lane takes).  The per-lane walk below lets every
// lane work through its own draw list, so with 32 lanes at different draws in different regimes the warp executes inversion
// at 7 lanes, the exact test at 8 lanes and the BTRS candidate at 26 on almost every trip -- 17.9 of 32 lanes per instruction,
// 423 thread-slots per variate of which 234 do work (profiles/r2_tau_leap_ncu_full.txt).  Round 2's first attempt (all lanes
// at the same draw, leftovers afterwards) raised the lanes per instruction but still executed the union of the regimes on
// every trip (121.7 against 108.8 ms).  This sampler separates the REGIMES instead of the draws:
//   * the (n, p) of all draws of the warp's trajectories sit in shared memory ([slot][lane]); a draw is an ITEM
//     (slot, lane, attempt number) of 16 bits and any lane may work on any item;
//   * classification (one uniform pass over the slots) settles the trivial draws and pushes the others on two stacks in the
//     warp's queue array: sequential inversion from the bottom, BTRS from the top (ballot + popc give the positions);
//   * the inversion stack is drained 32 items at a time -- every lane in the inversion loop;
//   * then BTRS: a batch pops 32 items, every lane draws a candidate and runs the box test -- all lanes in the same code.
//     Accepted candidates are stored; rejected ones go back on the BTRS stack with the next attempt number; undecided ones go
//     on the EXACT stack (bottom of the array, free by now).  When 32 undecided items have gathered (or nothing else is
//     left) a batch re-creates their candidates from the same words and runs the exact test on all lanes.
//   * draws conditional on another draw (DEP_MASK) form further LEVELS: level l starts when level l - 1 is final.
// With fewer than 32 trajectories per warp (small ensembles, lane_shift) the idle lanes work on their neighbours' items: a
// trajectory's 24 draws are spread over the warp.
// Random words are COUNTER-BASED per attempt: Philox2x32 keyed by the launch, counter = (trajectory, step, slot, attempt
// number), so what an item draws does not depend on which lane handles it, on its warp neighbours, on the launch geometry
// or on the number of GPUs.  Rejection sampling stays exact under any schedule: every attempt uses words nobody has looked at
// before, and whether it is accepted depends on its own words only.
// ------------------------------------------------------------------------------------------------------------------
#ifndef SODM_QUEUE_PHILOX_ROUNDS
#define SODM_QUEUE_PHILOX_ROUNDS 10
#endif

// Coordinates of an attempt -> Philox2x32 counter.  c0 = low 32 bits of the global trajectory id; c1 = step (16 bits) |
// slot (6) | attempt (6) | bits 32..35 of the trajectory id; steps beyond 2^16 move into the key (warp-uniform).
struct AttemptKey {
    uint32_t key;       // launch key (+ high bits of the step)
    uint32_t step16;    // (step & 0xffff) << 16
    uint64_t traj0;     // global id of the trajectory of lane 0 of this warp
    __host__ __device__ __forceinline__ void set(uint64_t seed, uint64_t traj_lane0, uint32_t step) {
        key = fold_seed(seed) + (step >> 16) * 0x9E3779B1u;
        step16 = step << 16;
        traj0 = traj_lane0;
    }
    __host__ __device__ __forceinline__ void words(int l, int d, uint32_t attempt, uint32_t& wu, uint32_t& wv) const {
        const uint64_t t = traj0 + (uint64_t)l;
        philox2x32<SODM_QUEUE_PHILOX_ROUNDS>(key, (uint32_t)t, step16 | ((uint32_t)d << 10) | (attempt << 4) | (uint32_t)((t >> 32) & 15u), wu, wv);
    }
};
constexpr int64_t MAX_TRAJ_ID = (int64_t)1 << 36;

// chain position of every slot: 0 for an independent draw, l for the l-th draw conditional on it (bit d of DEP_MASK)
template <int NDRAW, uint64_t DEP_MASK>
struct DrawLevels {
    int level[NDRAW] = {};
    int n_levels = 1;
    uint64_t mask[4] = {};          // slots of level 0..3
    constexpr DrawLevels() {
        for (int d = 0; d < NDRAW; ++d) {
            level[d] = ((DEP_MASK >> d) & 1ull) ? level[d - 1] + 1 : 0;
            if (level[d] + 1 > n_levels) n_levels = level[d] + 1;
            if (level[d] < 4) mask[level[d]] |= 1ull << d;
        }
    }
};

#ifdef __CUDA_ARCH__
__device__ __forceinline__ int coop_rank(unsigned m, int lane) { return __popc(m & ((1u << lane) - 1u)); }
__device__ __forceinline__ void coop_sync() { __syncwarp(); }
constexpr int COOP_WIDTH = 32;
#else
constexpr int COOP_WIDTH = 1;
inline int coop_rank(unsigned, int) { return 0; }                  // CPU harness: a "warp" of one lane runs the same algorithm
inline void coop_sync() {}
#endif

// item = slot << 5 | lane | attempt << 11
__host__ __device__ __forceinline__ uint32_t next_attempt(uint32_t item) {
    // 5 bits of attempt number; an item rejected 31 times in a row (probability < 1e-18) continues on numbers 16..31
    return (item >> 11) == 31u ? ((item & 0x7ffu) | (16u << 11)) : item + (1u << 11);
}

template <int NDRAW, uint64_t DEP_MASK>
__host__ __device__ __forceinline__ void sample_worklist_queue(const AttemptKey& key, WorkList& w, bool active, uint32_t& st) {
    static_assert(NDRAW <= 64, "at most 64 draws per step");
    constexpr DrawLevels<NDRAW, DEP_MASK> LV{};
    static_assert(LV.n_levels <= 4, "at most 3 conditional draws in a row");
    constexpr int CAP = NDRAW * COOP_WIDTH, WIDTH = COOP_WIDTH;
    const int S = w.stride, lane = w.lane;
    uint16_t* Q = w.queue;
#pragma unroll 1
    for (int lvl = 0; lvl < LV.n_levels; ++lvl) {
        const uint64_t slots = lvl == 0 ? LV.mask[0] : (lvl == 1 ? LV.mask[1] : (lvl == 2 ? LV.mask[2] : LV.mask[3]));
        // ---- first attempt in place: every lane takes slot d of its OWN trajectory -- the trajectories of an ensemble resemble
        // each other, so the lanes are almost always in the same regime.  Trivial draws and accepted BTRS candidates are
        // final; the rest becomes items on three stacks: inversion (bottom of the first half of the array, upwards), rejected
        // candidates (top of the first half, downwards, attempt 1), undecided candidates (second half).  An item is on one
        // stack at a time, so the first two cannot meet.
        int nI = 0, nB = 0, nE = 0;
        constexpr int MID = CAP;
#pragma unroll 1
        for (uint64_t todo = slots; todo != 0; todo &= todo - 1) {
            const int d = low_bit(todo);
            int state = BIN_DONE;
            bool retry = false, undecided = false;
            if (active) {
                int n = w.n[d * S + lane];
                const float p = w.p[d * S + lane];
                if (lvl > 0) {                                      // the earlier exits of this compartment element are final
                    n -= chain_taken<DEP_MASK>(w, d);
                    w.n[d * S + lane] = n;
                }
                int k;
                state = binomial_classify(n, p, k);
                if (state == BIN_BTRS) {
                    const bool flip = p > 0.5f;
                    const float pp = flip ? 1.0f - p : p;
                    uint32_t wu, wv;
                    key.words(lane, d, 0u, wu, wv);
                    int kk;
                    BtrsHat hat;
                    int s2 = btrs_candidate(n, pp, wu, wv, kk, hat);
                    if (s2 == BIN_SQUEEZE && n < 16777216 && (float)n * pp >= 0.9f * SODM_F32_MODE_LIMIT) {
                        float lhs;
                        s2 = btrs_squeeze(n, pp, kk, hat, lhs);
                    }
                    retry = s2 == BIN_RETRY;
                    undecided = s2 == BIN_SQUEEZE || s2 == BIN_EXACT;
                    k = flip ? n - kk : kk;
                    if (s2 == BIN_DONE) state = BIN_DONE;
                }
                if (state == BIN_DONE) {
                    if (k < 0) { st |= ST_BAD_PROB; k = 0; }
                    w.n[d * S + lane] = k;
                }
            }
            const unsigned mI = coop_ballot(state == BIN_INVERT), mR = coop_ballot(retry), mE = coop_ballot(undecided);
            const uint32_t item = (uint32_t)((d << 5) | lane);
            if (state == BIN_INVERT) Q[nI + coop_rank(mI, lane)] = (uint16_t)item;
            if (retry) Q[CAP - 1 - (nB + coop_rank(mR, lane))] = (uint16_t)(item | (1u << 11));
            if (undecided) Q[MID + nE + coop_rank(mE, lane)] = (uint16_t)item;
            nI += coop_popc(mI);
            nB += coop_popc(mR);
            nE += coop_popc(mE);
        }
        coop_sync();
        // ---- sequential inversion, 32 items at a time ------------------------------------------------------------------------
        for (int base = 0; base < nI; base += WIDTH) {
            if (base + lane < nI) {
                const uint32_t item = Q[base + lane];
                const int d = (int)(item >> 5) & 63, l = (int)item & 31;
                const int n = w.n[d * S + l];
                const float p = w.p[d * S + l];
                const bool flip = p > 0.5f;
                uint32_t wu, wv;
                key.words(l, d, 0u, wu, wv);
                const int kk = binomial_inversion_word(n, flip ? 1.0f - p : p, wu);
                w.n[d * S + l] = flip ? n - kk : kk;
            }
        }
        coop_sync();
        // ---- leftovers: rejected candidates and exact tests ---------------------
        while (nB > 0 || nE > 0) {
            const bool candidates = nB >= WIDTH || (nB > 0 && nE < WIDTH);
            const int have = candidates ? nB : nE;
            const int take = have < WIDTH ? have : WIDTH;
            const bool mine = lane < take;
            uint32_t item = 0u;
            if (mine) item = candidates ? Q[CAP - nB + lane] : Q[MID + nE - take + lane];
            if (candidates) nB -= take; else nE -= take;
            coop_sync();                                            // the popped entries have been read before anything is pushed
            bool retry = false, undecided = false;
            if (mine) {
                const int d = (int)(item >> 5) & 63, l = (int)item & 31;
                const int n = w.n[d * S + l];
                const float p = w.p[d * S + l];
                const bool flip = p > 0.5f;
                const float pp = flip ? 1.0f - p : p;
                uint32_t wu, wv;
                key.words(l, d, item >> 11, wu, wv);
                int kk;
                BtrsHat hat;
                int state = btrs_candidate(n, pp, wu, wv, kk, hat);
                if (candidates) {
                    // large modes, whose exact test is the FP64 routine, pass the squeeze first (it settles 98 % there)
                    if (state == BIN_SQUEEZE && n < 16777216 && (float)n * pp >= 0.9f * SODM_F32_MODE_LIMIT) {
                        float lhs;
                        state = btrs_squeeze(n, pp, kk, hat, lhs);
                    }
                    retry = state == BIN_RETRY;
                    undecided = state == BIN_SQUEEZE || state == BIN_EXACT;
                } else {
                    retry = !btrs_exact_accept(n, pp, kk, btrs_lhs(hat));
                }
                if (!retry && !undecided) w.n[d * S + l] = flip ? n - kk : kk;
            }
            const unsigned mR = coop_ballot(retry), mE = coop_ballot(undecided);
            if (retry) Q[CAP - 1 - (nB + coop_rank(mR, lane))] = (uint16_t)next_attempt(item);
            if (undecided) Q[MID + nE + coop_rank(mE, lane)] = (uint16_t)item;
            nB += coop_popc(mR);
            nE += coop_popc(mE);
            coop_sync();
        }
    }
}

// (Measured and dropped in round 2: POOLING the exact tests of a trip over the warp -- four lanes per candidate, each computing a
// quarter of the FP32 log-pmf ratio, candidates fetched by shuffle.  The exact test is a third of the walk's executed
// instructions at 8-11 active lanes, and the pooled form runs it at 32; but the warp vote that pooling needs makes all lanes
// meet on every trip, and the walk lives on lanes NOT meeting: a lane inside a 20-iteration inversion holds nobody up while its
// neighbours run ahead as a diverged group.  107.7 -> 153 ms per 2^20 trajectories; with __fns for the candidate-to-lane map
// 175 ms.)
// PER-LANE WALK (default sampler): each lane draws its slots one after the other with a private Philox4x32 stream, retrying
// a rejected attempt on its next trip through the loop while its neighbours move on to their next slot -- the warp runs for
// max_lane(sum of attempts) trips instead of sum_slot(max_lane attempts); the exact test runs inside the trip that needs it.
// Every trip consumes exactly two words (an inversion uses one and drops the other), and every lane starts a step on a fresh
// block, so all lanes refill their 4-word block on the same trips (letting a lane carry a half-used block across steps halves
// the active lanes per refill and costs 11 %, profiles/r1_sampler_ab.txt).
struct WalkRng {
    Philox4x32 gen;
    uint32_t w0, w1, w2, w3;
    bool second;
};

template <int NDRAW, uint64_t DEP_MASK>
__host__ __device__ __forceinline__ void sample_worklist_walk(WalkRng& rng, WorkList& w, uint32_t& st) {
    const int S = w.stride, lane = w.lane;
    int q = 0, chain = 0;                // chain = what the earlier exits of the same compartment element have taken
    // Every trip is one attempt that consumes exactly two Philox words, and every lane starts a step on a fresh block
    // (the half-used one of the previous step is dropped): the lanes of a warp then refill on the same trips, so the
    // generator -- a sixth of the kernel's instructions -- always runs with the whole warp.  Letting each lane carry
    // its leftover pair across steps halves the active lanes per refill and costs 11 % (profiles/r1_sampler_ab.txt).
    rng.second = false;
    while (q < NDRAW) {
        const int d = (int)w.ord[q * S + lane];
        int n = w.n[d * S + lane];
        const float p = w.p[d * S + lane];
        const bool dep = (DEP_MASK >> d) & 1ull;
        if (dep) n -= chain;
        if (!rng.second) {
            uint32_t b[4];
            rng.gen.next4_rounds<SODM_WALK_PHILOX_ROUNDS>(b);
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
// still take part in the warp's votes (they have no draws to make) but neither load nor store.
template <class M, typename OutT, bool ALIGNED, bool QUEUE>
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
    AttemptKey akey;                       // counter-based words of the queue sampler
    WalkRng wrng;                          // private Philox4x32 stream of the walk
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
        if constexpr (QUEUE) {
            akey.set(L.seed, (uint64_t)(L.traj_id0 + i - w.lane), (uint32_t)k);
            sample_worklist_queue<M::NDRAW, (uint64_t)M::DEP_MASK>(akey, w, active, st);
        } else {
            if (k == 0) reorder_draws<M::NDRAW, (uint64_t)M::DEP_MASK>(w);
            sample_worklist_walk<M::NDRAW, (uint64_t)M::DEP_MASK>(wrng, w, st);
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
                if (active) emit_interp<M, OutT>(L, i, j, L.t0 + (double)j, pin_in_loop(t), t_new, y_old, y);
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
    int tdpf_row = -1;                     // table row whose values are in par[TDPF_OFFSET ..] at the moment
    while (t < t_end) {
        if (events++ >= L.max_events) { st |= ST_MAX_STEPS; break; }
        if constexpr (M::TDPF_COUNT > 0) {
            // The reference calls the time-dependent parameter functions at every event time (base.py:376-380).  Functions the
            // host could only tabulate are piecewise constant on the OPEN intervals (t0 + k tau, t0 + (k+1) tau) (checked
            // there, base.py of this package; which side a boundary belongs to is immaterial, an event falls on one with
            // probability zero) -- except the first event, which IS evaluated at t0 itself: entry 0 of tdpf_index holds the
            // value at t0, entry 1 + k the value on interval k.
            int64_t k = t == L.t0 ? 0 : 1 + (int64_t)floor((t - L.t0) / L.tau);
            k = k < 0 ? 0 : (k >= L.n_steps ? (int64_t)L.n_steps - 1 : k);
            const int row = L.tdpf_index[k];
            if (row != tdpf_row) {
                tdpf_row = row;
#pragma unroll
                for (int e = 0; e < M::TDPF_COUNT; ++e) par[M::TDPF_OFFSET + e] = L.tdpf_table[(int64_t)row * M::TDPF_COUNT + e];
            }
        }
        double a[M::NCH];
        if constexpr (functor_uses_time<M>::value) M::propensities(y, par, a, t);
        else M::propensities(y, par, a);
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
            emit_interp<M, OutT>(L, i, j, L.t0 + (double)j, pin_in_loop(t), t_new, y_old, y);
            ++j;
        }
        t = t_new;
    }
    while (j < L.n_out) emit_counts<M, OutT>(L, i, j++, y);   // only after an aborted run: hold the last state
    if (L.status) L.status[i] = (int32_t)st;
}

}  // namespace sodm
