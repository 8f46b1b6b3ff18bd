// Validation of a sodm_ode_args request and its translation into the by-value kernel argument block.
// Host-only; shared by the CUDA entry point (sodm_ode.cu) and the test-only CPU harness (tests/hostcheck).
#pragma once
#include "common.cuh"
#include "jump_setup.cuh"   // jump_n_steps
#include "ode_run.cuh"

namespace sodm {

inline int32_t ode_n_kept(int32_t n_out, int32_t out_stride, int32_t out_last_only) {
    if (n_out < 1) return 0;
    if (out_last_only) return 1;
    if (out_stride < 1) return 0;
    return (n_out - 1) / out_stride + 1;
}

// `info` describes the model (from the registry, or a generated model's own constants)
inline int build_ode_launch_for(const sodm_ode_args* a, OdeLaunch& L, const sodm_model_info& info) {
    if (!a) return fail(SODM_EINVAL, "sodm_ode_run: null args");
    if (info.kind != SODM_KIND_ODE) return fail(SODM_EMODEL, "sodm_ode_run: model %d is not an ODE model", a->model);
    if (a->n_traj < 0 || a->ld_in < a->n_traj || a->ld_out < a->n_traj)
        return fail(SODM_EINVAL, "sodm_ode_run: need 0 <= n_traj <= ld_in, ld_out");
    if (a->n_out < 1) return fail(SODM_EINVAL, "sodm_ode_run: n_out must be >= 1");
    if (a->n_traj == 0) { L = OdeLaunch{}; return SODM_OK; }                      // empty launch: buffers may be null
    if (!a->y0 || !a->par || !a->out) return fail(SODM_EINVAL, "sodm_ode_run: y0, par and out are required");
    if (a->n_var < 0 || a->n_var > info.n_param || (a->n_var && (!a->var_index || !a->var_values)))
        return fail(SODM_EINVAL, "sodm_ode_run: inconsistent per-trajectory parameter description");
    if (a->out_dtype != SODM_F32 && a->out_dtype != SODM_F64)
        return fail(SODM_EINVAL, "sodm_ode_run: out_dtype must be SODM_F32 or SODM_F64");
    if (a->method < SODM_ODE_RK45 || a->method > SODM_ODE_BDF)
        return fail(SODM_EINVAL, "sodm_ode_run: unknown method %d", a->method);
    L = OdeLaunch{};
    L.n_traj = a->n_traj;
    L.ld_in = a->ld_in;
    L.ld_out = a->ld_out;
    L.y0 = a->y0;
    L.y0_bcast = a->y0_broadcast;
    L.nx = a->model_arg;
    for (int e = 0; e < ODE_MAX_PARAM; ++e) { L.par[e] = e < info.n_param ? a->par[e] : 0.0; L.var_row[e] = -1; }
    for (int v = 0; v < a->n_var; ++v) {
        const int e = a->var_index[v];
        if (e < 0 || e >= info.n_param) return fail(SODM_EINVAL, "sodm_ode_run: var_index[%d]=%d out of range", v, e);
        L.var_row[e] = (int8_t)v;
    }
    L.var_values = a->var_values;
    L.t0 = a->t0;
    L.n_out = a->n_out;
    L.out_stride = a->out_stride < 1 ? 1 : a->out_stride;
    L.out_last_only = a->out_last_only ? 1 : 0;
    L.n_kept = ode_n_kept(L.n_out, L.out_stride, L.out_last_only);
    L.rtol = a->rtol;
    L.atol = a->atol;
    L.tau = a->tau;
    L.out = a->out;
    L.nfev = a->nfev;
    L.status = a->status;
    L.max_steps = a->max_steps > 0 ? a->max_steps : 2147483647LL;
    if (a->method == SODM_ODE_RK45 || a->method == SODM_ODE_RK23 || a->method >= SODM_ODE_DOP853) {
        if (!(a->rtol > 0.0) || !(a->atol >= 0.0)) return fail(SODM_EINVAL, "sodm_ode_run: need rtol > 0 and atol >= 0");
    } else {
        const int64_t n = jump_n_steps(a->t0, a->n_out, a->tau);
        if (n < 0) return fail(SODM_EINVAL, "sodm_ode_run: tau must be > 0 (and the run below 2^31 steps)");
        L.n_steps_fixed = (int32_t)n;
    }
    return SODM_OK;
}

#ifndef SODM_PLUGIN_BUILD
inline int build_ode_launch(const sodm_ode_args* a, OdeLaunch& L, sodm_model_info& info) {
    if (!a) return fail(SODM_EINVAL, "sodm_ode_run: null args");
    if (sodm_model_info_get(a->model, a->model_arg, &info) != SODM_OK) return SODM_EMODEL;
    return build_ode_launch_for(a, L, info);
}
#endif

}  // namespace sodm
