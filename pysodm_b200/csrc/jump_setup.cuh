// Validation of a sodm_jump_args request and its translation into the by-value kernel argument block.
// Host-only; shared by the CUDA entry point (sodm_jump.cu) and the test-only CPU harness
// (tests/hostcheck), which runs the same __host__ __device__ trajectory code without a GPU.
#pragma once
#include <cmath>
#include "common.cuh"
#include "jump_run.cuh"

namespace sodm {

// tau = 2^-k (k >= 0): step times are exact in FP64 and every integer output time is a step time
inline bool tau_is_aligned(double tau, int* steps_per_out) {
    if (!(tau > 0.0) || tau > 1.0) return false;
    const double inv = 1.0 / tau;
    if (inv != floor(inv) || inv > 1048576.0) return false;
    const int64_t n = (int64_t)inv;
    if ((n & (n - 1)) != 0) return false;
    *steps_per_out = (int)n;
    return true;
}


inline int64_t jump_n_steps(double t0, int32_t n_out, double tau) {
    if (!(tau > 0.0) || n_out < 1) return -1;
    const double t_end = t0 + (double)(n_out - 1);
    int64_t n = 0;
    double t = t0;                       // same FP64 accumulation as base.py:428-434
    while (t < t_end) {
        t = t + tau;
        if (++n > (int64_t)1 << 31) return -1;
    }
    return n;
}

// Fills L and `aligned`; returns SODM_OK or an error code (message in the thread-local buffer).
// `info` describes the model (from the registry, or a generated model's own constants)
inline int build_jump_launch_for(const sodm_jump_args* a, JumpLaunch& L, bool& aligned, const sodm_model_info& info) {
    if (!a) return fail(SODM_EINVAL, "sodm_jump_run: null args");
    if (info.kind != SODM_KIND_JUMP)
        return fail(SODM_EMODEL, "sodm_jump_run: model %d is not a jump-process model", a->model);
    if (a->n_traj < 0 || a->ld_in < a->n_traj || a->ld_out < a->n_traj)
        return fail(SODM_EINVAL, "sodm_jump_run: need 0 <= n_traj <= ld_in, ld_out");
    if (a->n_out < 1) return fail(SODM_EINVAL, "sodm_jump_run: n_out must be >= 1");
    if (a->n_traj == 0) { L = JumpLaunch{}; aligned = false; return SODM_OK; }   // empty launch: buffers may be null
    if (!a->y0 || !a->par || !a->out) return fail(SODM_EINVAL, "sodm_jump_run: y0, par and out are required");
    if (a->n_var < 0 || a->n_var > info.n_param || (a->n_var && (!a->var_index || !a->var_values)))
        return fail(SODM_EINVAL, "sodm_jump_run: inconsistent per-trajectory parameter description");
    if (a->method != SODM_JUMP_TAU_LEAP && a->method != SODM_JUMP_SSA)
        return fail(SODM_EINVAL, "sodm_jump_run: unknown method %d", a->method);

    L = JumpLaunch{};
    L.n_traj = a->n_traj;
    L.traj_id0 = a->traj_id0;
    L.ld_in = a->ld_in;
    L.ld_out = a->ld_out;
    L.seed = a->seed;
    L.y0 = a->y0;
    L.y0_bcast = a->y0_broadcast;
    for (int e = 0; e < MAX_PARAM; ++e) { L.par[e] = e < info.n_param ? a->par[e] : 0.0; L.var_row[e] = -1; }
    for (int v = 0; v < a->n_var; ++v) {
        const int e = a->var_index[v];
        if (e < 0 || e >= info.n_param) return fail(SODM_EINVAL, "sodm_jump_run: var_index[%d]=%d out of range", v, e);
        if (info.tdpf_count && e >= info.tdpf_offset && e < info.tdpf_offset + info.tdpf_count)
            return fail(SODM_EUNSUPPORTED, "sodm_jump_run: parameter element %d is time dependent and cannot also vary per trajectory", e);
        L.var_row[e] = (int8_t)v;
    }
    L.var_values = a->var_values;
    L.t0 = a->t0;
    L.tau = a->tau;
    L.n_out = a->n_out;
    for (int s = 0; s < MAX_STATE; ++s) L.out_row[s] = -1;
    if (a->n_sel == 0) {
        for (int s = 0; s < info.n_state; ++s) L.out_row[s] = (int8_t)s;
    } else {
        if (a->n_sel < 0 || a->n_sel > info.n_state || !a->state_sel)
            return fail(SODM_EINVAL, "sodm_jump_run: bad state selection");
        for (int r = 0; r < a->n_sel; ++r) {
            const int s = a->state_sel[r];
            if (s < 0 || s >= info.n_state || L.out_row[s] >= 0)
                return fail(SODM_EINVAL, "sodm_jump_run: state_sel[%d]=%d invalid or repeated", r, s);
            L.out_row[s] = (int8_t)r;
        }
    }
    for (int s = 0; s < MAX_STATE; ++s)
        L.out_off[s] = L.out_row[s] >= 0 ? (int64_t)L.out_row[s] * a->n_out * a->ld_out : -1;
    L.out = a->out;
    L.status = a->status;
    L.compat = a->compat;
    L.max_events = a->max_events > 0 ? a->max_events : ((int64_t)1 << 62);
    if (a->n_out_peers < 0 || a->n_out_peers > SODM_MAX_PEERS || (a->n_out_peers && !a->out_peers))
        return fail(SODM_EINVAL, "sodm_jump_run: 0 <= n_out_peers <= %d and out_peers must be given", SODM_MAX_PEERS);
    L.n_peers = a->n_out_peers;
    for (int p = 0; p < a->n_out_peers; ++p) {
        if (!a->out_peers[p]) return fail(SODM_EINVAL, "sodm_jump_run: out_peers[%d] is null", p);
        L.peer_out[p] = a->out_peers[p];
    }

    aligned = false;
    if (a->method == SODM_JUMP_TAU_LEAP) {
        const int64_t n_steps = jump_n_steps(a->t0, a->n_out, a->tau);
        if (n_steps < 0) return fail(SODM_EINVAL, "sodm_jump_run: tau must be > 0 (and the run below 2^31 steps)");
        L.n_steps = (int32_t)n_steps;
        if (a->traj_id0 < 0 || a->traj_id0 + a->n_traj > MAX_TRAJ_ID)
            return fail(SODM_EINVAL, "sodm_jump_run: tau-leap trajectory ids must lie in [0, 2^36) (counter layout of the sampler's random words)");
        int spo = 0;
        aligned = tau_is_aligned(a->tau, &spo);
        L.steps_per_out = aligned ? spo : 0;
        if (info.tdpf_count) {
            if (!a->tdpf_table || !a->tdpf_index || a->n_tdpf_values < 1 || a->n_tdpf_index < n_steps)
                return fail(SODM_EINVAL, "sodm_jump_run: model needs a time-dependent parameter table with >= %lld step entries",
                            (long long)n_steps);
            if ((size_t)a->n_tdpf_values * info.tdpf_count * sizeof(float) > 16 * 1024)
                return fail(SODM_EUNSUPPORTED, "sodm_jump_run: more than %d distinct time-dependent values",
                            (int)(16 * 1024 / sizeof(float) / info.tdpf_count));
            L.n_tdpf_values = a->n_tdpf_values;
            L.tdpf_index = a->tdpf_index;
            L.tdpf_table = a->tdpf_table;
        }
    } else {
        if (!info.supports_ssa) return fail(SODM_EUNSUPPORTED, "sodm_jump_run: model %d has no SSA functor", a->model);
        if (info.tdpf_count) {
            // SSA evaluates time-dependent parameters at event times: tdpf_index[0] = row in force at t0 itself (the first
            // event), tdpf_index[1 + k] = row in force on the open interval (t0 + k tau, t0 + (k+1) tau)
            if (!a->tdpf_table || !a->tdpf_index || a->n_tdpf_values < 1 || a->n_tdpf_index < 1 || !(a->tau > 0.0))
                return fail(SODM_EINVAL, "sodm_jump_run: SSA of a model with time-dependent parameters needs their table, "
                                         "an interval index and the interval width in `tau`");
            L.n_tdpf_values = a->n_tdpf_values;
            L.tdpf_index = a->tdpf_index;
            L.tdpf_table = a->tdpf_table;
            L.n_steps = a->n_tdpf_index;
        }
    }
    if (a->out_dtype == SODM_I32 && !aligned)
        return fail(SODM_EINVAL, "sodm_jump_run: int32 output needs tau-leap with tau = 2^-k (outputs are interpolated otherwise)");

    return SODM_OK;
}

#ifndef SODM_PLUGIN_BUILD
inline int build_jump_launch(const sodm_jump_args* a, JumpLaunch& L, bool& aligned) {
    if (!a) return fail(SODM_EINVAL, "sodm_jump_run: null args");
    sodm_model_info info;
    if (sodm_model_info_get(a->model, 0, &info) != SODM_OK)
        return fail(SODM_EMODEL, "sodm_jump_run: model %d is not registered", a->model);
    return build_jump_launch_for(a, L, aligned, info);
}
#endif

}  // namespace sodm
