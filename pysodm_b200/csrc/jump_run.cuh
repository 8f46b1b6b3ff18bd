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

constexpr int MAX_PARAM = 32;   // flat parameter elements passed by value
constexpr int MAX_STATE = 32;   // flat state elements of a jump model

enum : uint32_t { COMPAT_STAY_QUIRK = 1u, COMPAT_SSA_FIRST_MATCH = 2u, COMPAT_WALK_DECOUPLED = 256u, COMPAT_WALK_LOCKSTEP = 512u };

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
    int32_t* status;           // [n_traj] or null
    uint32_t compat;
    int64_t max_events;        // SSA event budget per trajectory
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
    OutT* out = (OutT*)L.out + i;
#pragma unroll
    for (int s = 0; s < M::NS; ++s) {
        if (L.out_row[s] >= 0) out[((int64_t)L.out_row[s] * L.n_out + j) * L.ld_out] = (OutT)y[s];
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

// The binomial draws of one step of one trajectory: slot d holds (n_d, p_d) and, once drawn, T_d in
// place of n_d.  On the device the slots live in shared memory, one column per thread (element
// [d][thread]), so lanes that are at DIFFERENT slots still hit 32 different banks.
struct WorkList {
    int* n;
    float* p;
    uint32_t* a;        // phased sampler scratch: random word / BTRS candidate of a pending slot
    uint32_t* b;        //                         random word / log of the scaled uniform of a pending slot
    int stride;
    unsigned lanes;     // device: lanes of this warp that run a trajectory (mask for __syncwarp)
    __host__ __device__ __forceinline__ void put(int d, int nd, float pd) { n[d * stride] = nd; p[d * stride] = pd; }
    __host__ __device__ __forceinline__ int get(int d) const { return n[d * stride]; }
};

__host__ __device__ __forceinline__ int lowest_bit(uint32_t m) {
#ifdef __CUDA_ARCH__
    return __ffs((int)m) - 1;
#else
    return __builtin_ctz(m);
#endif
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

// PHASED sampler (default).  The cost of a divergent SIMT loop is (trips) x (union of the branches any lane takes), so
// a step's draws are organised in phases whose bodies are homogeneous:
//   A  first pass, uniform over the warp: one Philox block per two slots; every slot is classified (trivial /
//      inversion regime / BTRS regime) and BTRS slots get their first attempt.  ~80 % of them finish here.
//   X  the candidates the squeeze could not settle: exact log-pmf test, lanes pop their own pending slots;
//   R  rejected slots: fresh attempts from a second Philox stream until accepted (lanes pop their own slots);
//   I  inversion-regime slots: sequential search, lanes pop their own slots.
// Slots conditional on another slot (DEP_MASK) run as a second level after their predecessors are final; their random
// words are drawn in the first level's pass A and parked in the scratch arrays.
template <int NDRAW, uint32_t DEP_MASK, class Rng>
__host__ __device__ __forceinline__ void sample_worklist_phased(Philox4x32& first, Rng& retry, WorkList& w, uint32_t& st) {
    static_assert(NDRAW <= 32, "at most 32 draws per step");
    constexpr int LEVELS = max_level(DEP_MASK, NDRAW) + 1;
    const int S = w.stride;
#pragma unroll
    for (int level = 0; level < LEVELS; ++level) {
        const uint32_t in_level = level_mask(DEP_MASK, NDRAW, level);
        uint32_t inv = 0u, exact = 0u, again = 0u;
        // ---- phase A --------------------------------------------------------------------------------------
        uint32_t blk[4] = {0u, 0u, 0u, 0u};
#pragma unroll 1
        for (int d = 0; d < NDRAW; ++d) {
            uint32_t wu, wv;
            if (level == 0) {
                if ((d & 1) == 0) first.next4(blk);
                wu = (d & 1) ? blk[2] : blk[0];
                wv = (d & 1) ? blk[3] : blk[1];
                if (!((in_level >> d) & 1u)) {            // a later level's slot: park its words
                    w.a[d * S] = wu;
                    w.b[d * S] = wv;
                    continue;
                }
            } else {
                if (!((in_level >> d) & 1u)) continue;
                wu = w.a[d * S];
                wv = w.b[d * S];
            }
            int n = w.n[d * S];
            const float p = w.p[d * S];
            if (level > 0) { n -= w.n[(d - 1) * S]; w.n[d * S] = n; }     // what the previous exit left behind
            int k;
            int state = binomial_classify(n, p, k);
            if (state == BIN_BTRS) {
                const bool flip = p > 0.5f;
                float lhs = 0.0f;
                state = btrs_attempt_words(n, flip ? 1.0f - p : p, wu, wv, k, lhs);
                if (state == BIN_DONE) k = flip ? n - k : k;
                else if (state == BIN_EXACT) { w.a[d * S] = (uint32_t)k; w.b[d * S] = float_as_word(lhs); exact |= 1u << d; }
                else again |= 1u << d;
            } else if (state == BIN_INVERT) {
                w.a[d * S] = wu;
                inv |= 1u << d;
            }
            if (state == BIN_DONE) {
                if (k < 0) { st |= ST_BAD_PROB; k = 0; }
                w.n[d * S] = k;
            }
        }
        // ---- phases X + R: settle ambiguous candidates, re-draw rejected ones ------------------------------
        while (exact | again) {
            if (exact) {
                const int d = lowest_bit(exact);
                const uint32_t bit = 1u << d;
                const int n = w.n[d * S];
                const float p = w.p[d * S];
                const bool flip = p > 0.5f;
                const int k = (int)w.a[d * S];
                exact &= ~bit;
                if (btrs_exact_accept(n, flip ? 1.0f - p : p, k, word_as_float(w.b[d * S]))) w.n[d * S] = flip ? n - k : k;
                else again |= bit;
            }
            if (again) {
                const int d = lowest_bit(again);
                const uint32_t bit = 1u << d;
                const int n = w.n[d * S];
                const float p = w.p[d * S];
                const bool flip = p > 0.5f;
                int k;
                float lhs = 0.0f;
                const uint32_t wu = retry.next(), wv = retry.next();
                const int state = btrs_attempt_words(n, flip ? 1.0f - p : p, wu, wv, k, lhs);
                if (state == BIN_DONE) { w.n[d * S] = flip ? n - k : k; again &= ~bit; }
                else if (state == BIN_EXACT) { w.a[d * S] = (uint32_t)k; w.b[d * S] = float_as_word(lhs); exact |= bit; again &= ~bit; }
            }
        }
        // ---- phase I: inversion-regime slots ----------------------------------------------------------------
        while (inv) {
            const int d = lowest_bit(inv);
            inv &= inv - 1u;
            const int n = w.n[d * S];
            const float p = w.p[d * S];
            const bool flip = p > 0.5f;
            const int k = binomial_inversion_word(n, flip ? 1.0f - p : p, w.a[d * S]);
            w.n[d * S] = flip ? n - k : k;
        }
#ifdef __CUDA_ARCH__
        __syncwarp(w.lanes);
#endif
    }
}

// Draws every slot of the list.
//   lockstep (default): all lanes of the warp are at the SAME slot -- i.e. the same compartment and age group of
//     32 neighbouring trajectories, whose (n, p) are alike, so the lanes take the same sampler branch; a lane whose
//     attempt is rejected retries while its neighbours wait (max over lanes of a geometric number of attempts).
//   decoupled (COMPAT_WALK_DECOUPLED): each lane walks its OWN list and never waits for a neighbour's rejection,
//     at the price of lanes drifting to different slots and hence different sampler branches.
// Which one wins is a measurement (profiles/): on influenza the lockstep walk executes ~5x fewer warp instructions.
template <int NDRAW, uint32_t DEP_MASK, class Rng>
__host__ __device__ __forceinline__ void sample_worklist_walk(Rng& rng, WorkList& w, uint32_t& st, bool decoupled) {
    int prev = 0;
    if (!decoupled) {
#pragma unroll 1
        for (int d = 0; d < NDRAW; ++d) {
            int n = w.n[d * w.stride];
            const float p = w.p[d * w.stride];
            if ((DEP_MASK >> d) & 1u) n -= prev;      // conditional draw: what the previous exit left behind
            int k;
            while (!binomial_try(rng, n, p, k)) {}
#ifdef __CUDA_ARCH__
            __syncwarp(w.lanes);                      // reconverge before the next slot
#endif
            if (k < 0) { st |= ST_BAD_PROB; k = 0; }
            w.n[d * w.stride] = k;
            prev = k;
        }
        return;
    }
    int d = 0;
    while (d < NDRAW) {
        int n = w.n[d * w.stride];
        const float p = w.p[d * w.stride];
        if ((DEP_MASK >> d) & 1u) n -= prev;
        int k;
        if (binomial_try(rng, n, p, k)) {
            if (k < 0) { st |= ST_BAD_PROB; k = 0; }
            w.n[d * w.stride] = k;
            prev = k;
            ++d;
        }
    }
}

template <class M, typename OutT, bool ALIGNED>
__host__ __device__ __forceinline__ void tau_leap_trajectory(const JumpLaunch& L, int64_t i, const float* table,
                                                             WorkList w) {
    double par[M::NPAR];
    int y[M::NS];
    uint32_t st = ST_OK;
    load_trajectory<M>(L, i, par, y, st);
    const bool quirk = (L.compat & COMPAT_STAY_QUIRK) != 0u;
    const uint32_t walk = L.compat & (COMPAT_WALK_DECOUPLED | COMPAT_WALK_LOCKSTEP);
    const float tau = (float)L.tau;
    const typename M::Pre pre = M::prepare(par, tau, quirk);
    PhiloxWords rng;                       // stream 0: re-draws (phased sampler) / every draw (walk samplers)
    rng.init(L.seed, (uint64_t)(L.traj_id0 + i));
    Philox4x32 first;                      // stream 2: first-attempt words of the phased sampler
    first.init(L.seed, (uint64_t)(L.traj_id0 + i), 2u);

    emit_counts<M, OutT>(L, i, 0, y);
    int j = 1;
    if (ALIGNED) {
        int until_out = L.steps_per_out;
        for (int k = 0; k < L.n_steps; ++k) {
            const float* tab = M::TDPF_COUNT ? table + (int64_t)L.tdpf_index[k] * M::TDPF_COUNT : table;
            M::propose(y, pre, tab, tau, quirk, w);
            if (walk == 0u) sample_worklist_phased<M::NDRAW, M::DEP_MASK>(first, rng, w, st);
            else sample_worklist_walk<M::NDRAW, M::DEP_MASK>(rng, w, st, walk == COMPAT_WALK_DECOUPLED);
            M::apply(y, w);
            if (--until_out == 0) {
                emit_counts<M, OutT>(L, i, j++, y);
                until_out = L.steps_per_out;
            }
        }
    } else {
        // general tau: interpolate onto the integer grid between consecutive step states
        int y_old[M::NS];
        double t = L.t0;
        for (int k = 0; k < L.n_steps; ++k) {
#pragma unroll
            for (int s = 0; s < M::NS; ++s) y_old[s] = y[s];
            const float* tab = M::TDPF_COUNT ? table + (int64_t)L.tdpf_index[k] * M::TDPF_COUNT : table;
            M::propose(y, pre, tab, tau, quirk, w);
            if (walk == 0u) sample_worklist_phased<M::NDRAW, M::DEP_MASK>(first, rng, w, st);
            else sample_worklist_walk<M::NDRAW, M::DEP_MASK>(rng, w, st, walk == COMPAT_WALK_DECOUPLED);
            M::apply(y, w);
            const double t_new = t + L.tau;
            while (j < L.n_out && L.t0 + (double)j <= t_new) {
                emit_interp<M, OutT>(L, i, j, L.t0 + (double)j, t, t_new, y_old, y);
                ++j;
            }
            t = t_new;
        }
    }
    if (L.status) L.status[i] = (int32_t)st;
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
