// Entry points of a model plugin (include/sodm_plugin.h): a generated translation unit defines ONE functor and expands
// SODM_JUMP_PLUGIN(Functor) or SODM_ODE_PLUGIN(Functor); everything else -- kernels, launch validation, samplers,
// integrators -- is the code of the built-in models, instantiated for the new functor.
#pragma once
#define SODM_PLUGIN_BUILD 1
#include "../../include/sodm_plugin.h"
#include "jump_kernels.cuh"
#include "ode_kernels.cuh"

namespace sodm {
char* error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}
}  // namespace sodm

#define SODM_PLUGIN_COMMON()                                                                      \
    extern "C" int sodm_plugin_abi_version(void) { return SODM_ABI_VERSION; }                    \
    extern "C" const char* sodm_plugin_last_error(void) { return ::sodm::error_buffer(); }

#define SODM_JUMP_PLUGIN(M)                                                                       \
    SODM_PLUGIN_COMMON()                                                                          \
    static sodm_model_info plugin_info() {                                                        \
        sodm_model_info info = {};                                                                \
        info.kind = SODM_KIND_JUMP;                                                               \
        info.n_state = M::NS;                                                                     \
        info.n_param = M::NPAR;                                                                   \
        info.n_channel = M::NCH;                                                                  \
        info.tdpf_offset = M::TDPF_OFFSET;                                                        \
        info.tdpf_count = M::TDPF_COUNT;                                                          \
        info.supports_ssa = 1;                                                                    \
        return info;                                                                              \
    }                                                                                             \
    extern "C" int sodm_plugin_info(sodm_model_info* out) {                                       \
        if (!out) return ::sodm::fail(SODM_EINVAL, "sodm_plugin_info: null output");              \
        *out = plugin_info();                                                                     \
        return SODM_OK;                                                                           \
    }                                                                                             \
    extern "C" int sodm_plugin_jump_run(const sodm_jump_args* a) {                                \
        static_assert(M::NS <= ::sodm::MAX_STATE && M::NPAR <= ::sodm::MAX_PARAM, "model too large for one launch block"); \
        ::sodm::JumpLaunch L;                                                                     \
        bool aligned = false;                                                                     \
        const int rc = ::sodm::build_jump_launch_for(a, L, aligned, plugin_info());               \
        if (rc != SODM_OK) return rc;                                                             \
        if (a->n_traj == 0) return SODM_OK;                                                       \
        if (a->compat & ::sodm::TUNE_QUEUE)                                                 \
            return ::sodm::fail(SODM_EUNSUPPORTED, "generated models use the default sampler only"); \
        return ::sodm::launch_plugin_jump<M>(L, a->method, aligned, a->out_dtype, (cudaStream_t)a->stream); \
    }                                                                                             \
    extern "C" int sodm_plugin_ode_run(const sodm_ode_args*) {                                    \
        return ::sodm::fail(SODM_EUNSUPPORTED, "this plugin holds a jump-process model");         \
    }

#define SODM_ODE_PLUGIN(M)                                                                        \
    SODM_PLUGIN_COMMON()                                                                          \
    static sodm_model_info plugin_info() {                                                        \
        sodm_model_info info = {};                                                                \
        info.kind = SODM_KIND_ODE;                                                                \
        info.n_state = M::N;                                                                      \
        info.n_param = M::NPAR;                                                                   \
        return info;                                                                              \
    }                                                                                             \
    extern "C" int sodm_plugin_info(sodm_model_info* out) {                                       \
        if (!out) return ::sodm::fail(SODM_EINVAL, "sodm_plugin_info: null output");              \
        *out = plugin_info();                                                                     \
        return SODM_OK;                                                                           \
    }                                                                                             \
    extern "C" int sodm_plugin_ode_run(const sodm_ode_args* a) {                                  \
        static_assert(M::NPAR <= ::sodm::ODE_MAX_PARAM, "too many flat parameter elements");      \
        ::sodm::OdeLaunch L;                                                                      \
        const int rc = ::sodm::build_ode_launch_for(a, L, plugin_info());                         \
        if (rc != SODM_OK) return rc;                                                             \
        if (a->n_traj == 0) return SODM_OK;                                                       \
        return ::sodm::launch_thread<M>(L, a->method, a->out_dtype, (cudaStream_t)a->stream);     \
    }                                                                                             \
    extern "C" int sodm_plugin_jump_run(const sodm_jump_args*) {                                  \
        return ::sodm::fail(SODM_EUNSUPPORTED, "this plugin holds an ODE model");                 \
    }

// Large generated ODE systems: warp-per-draw (codegen.ode_warp_source).  W = the generated shape functor (GenWarp), M = the
// thread functor of the same model, kept for the host self-check only.
#define SODM_ODE_WARP_PLUGIN(M, W)                                                                \
    SODM_PLUGIN_COMMON()                                                                          \
    static sodm_model_info plugin_info() {                                                        \
        sodm_model_info info = {};                                                                \
        info.kind = SODM_KIND_ODE;                                                                \
        info.n_state = W::N;                                                                      \
        info.n_param = W::NPAR;                                                                   \
        return info;                                                                              \
    }                                                                                             \
    extern "C" int sodm_plugin_info(sodm_model_info* out) {                                       \
        if (!out) return ::sodm::fail(SODM_EINVAL, "sodm_plugin_info: null output");              \
        *out = plugin_info();                                                                     \
        return SODM_OK;                                                                           \
    }                                                                                             \
    extern "C" int sodm_plugin_ode_run(const sodm_ode_args* a) {                                  \
        static_assert(W::NPAR <= ::sodm::ODE_MAX_PARAM, "too many flat parameter elements");      \
        ::sodm::OdeLaunch L;                                                                      \
        const int rc = ::sodm::build_ode_launch_for(a, L, plugin_info());                         \
        if (rc != SODM_OK) return rc;                                                             \
        if (a->n_traj == 0) return SODM_OK;                                                       \
        return ::sodm::launch_gen_warp<W>(L, a->method, a->out_dtype, (cudaStream_t)a->stream);   \
    }                                                                                             \
    extern "C" int sodm_plugin_jump_run(const sodm_jump_args*) {                                  \
        return ::sodm::fail(SODM_EUNSUPPORTED, "this plugin holds an ODE model");                 \
    }

namespace sodm {

// One warp per draw for a generated ODE: the draw's state and parameters live in shared memory between stages; position k
// of lane l holds equation W::flat_index(k, l) (equations grouped by expression shape, groups padded to whole warps), so
// that `rhs_lane` runs ONE template per k for the whole warp.
template <class W>
struct GenWarpSystem {
    static constexpr int NL = W::NL;
    double* ys;            // shared: full state of this draw (input of the current right-hand-side evaluation)
    const double* pr;      // shared: parameters of this draw
    int lane;
    uint32_t valid;        // bit k: position k of this lane holds an equation
    __device__ __forceinline__ int n_total() const { return W::N; }
    __device__ __forceinline__ bool active(int k) const { return (valid >> k) & 1u; }
    __device__ __forceinline__ bool leader() const { return lane == 0; }
    __device__ __forceinline__ int global_index(int k) const { return W::flat_index(k, lane); }
    __device__ __forceinline__ double reduce(double x) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        return x;
    }
    __device__ __forceinline__ void rhs(double t, const double* y, double* dy) const {
        __syncwarp();                                            // every lane is done reading the previous stage's state
#pragma unroll
        for (int k = 0; k < NL; ++k)
            if (active(k)) ys[W::flat_index(k, lane)] = y[k];
        __syncwarp();
        W::rhs_lane(t, ys, pr, lane, dy);
    }
};

#ifndef SODM_GEN_WARP_MIN_BLOCKS
#define SODM_GEN_WARP_MIN_BLOCKS 3       // as for the hand-written reactor kernel: FP64 dependency chains want resident warps
#endif
template <class W, typename OutT>
__global__ void __launch_bounds__(ODE_BLOCK, SODM_GEN_WARP_MIN_BLOCKS) ode_gen_warp_kernel(const __grid_constant__ OdeLaunch L, int method) {
    extern __shared__ double gen_warp_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= L.n_traj) return;                                   // whole warps leave together
    double* ys = gen_warp_smem + (size_t)warp * (W::N + W::NPAR);
    double* pr = ys + W::N;
    for (int e = lane; e < W::NPAR; e += 32)
        pr[e] = L.var_row[e] >= 0 ? L.var_values[(int64_t)L.var_row[e] * L.ld_in + i] : L.par[e];
    GenWarpSystem<W> sys{ys, pr, lane, 0u};
    double y[W::NL];
#pragma unroll
    for (int k = 0; k < W::NL; ++k) {
        const int j = W::flat_index(k, lane);
        y[k] = 0.0;
        if (j >= 0) {
            sys.valid |= 1u << k;
            y[k] = L.y0_bcast ? L.y0[j] : L.y0[(int64_t)j * L.ld_in + i];
        }
    }
    __syncwarp();
    if (method == SODM_ODE_RK45) rk45_run<GenWarpSystem<W>, OutT>(sys, L, i, y);
    else if (method == SODM_ODE_RK23) rk45_run<GenWarpSystem<W>, OutT, SchemeRK23>(sys, L, i, y);
    else if (method == SODM_ODE_DOP853) return;                  // refused by the launcher: 16 stage vectors x NL positions per lane
    else if (method == SODM_ODE_EULER) fixed_step_run<GenWarpSystem<W>, OutT, false>(sys, L, i, y);
    else fixed_step_run<GenWarpSystem<W>, OutT, true>(sys, L, i, y);
}

template <class W>
static int launch_gen_warp(const OdeLaunch& L, int method, int dtype, cudaStream_t stream) {
    static_assert(W::NL <= 32, "at most 32 positions per lane");
    if (method == SODM_ODE_RADAU || method == SODM_ODE_BDF)
        return fail(SODM_EUNSUPPORTED, "Radau / BDF are offered for thread-per-trajectory models (at most %d equations), not for the warp-per-draw layout", IMPLICIT_MAX_N);
    if (method == SODM_ODE_DOP853)
        return fail(SODM_EUNSUPPORTED, "DOP853 is not offered for generated models in the warp-per-draw layout (16 stage vectors per lane)");
    const int64_t threads = L.n_traj * 32;
    const unsigned grid = (unsigned)((threads + ODE_BLOCK - 1) / ODE_BLOCK);
    const size_t smem = (size_t)(ODE_BLOCK / 32) * (W::N + W::NPAR) * sizeof(double);
    if (smem > 200 * 1024) return fail(SODM_EUNSUPPORTED, "model needs %zu bytes of shared memory per block", smem);
    void (*kernel)(OdeLaunch, int) = dtype == SODM_F64 ? ode_gen_warp_kernel<W, double> : ode_gen_warp_kernel<W, float>;
    SODM_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 48 * 1024 ? smem : 48 * 1024)));
    kernel<<<grid, ODE_BLOCK, smem, stream>>>(L, method);
    SODM_CUDA_CHECK(cudaGetLastError());
    return SODM_OK;
}

// A generated model gets the kernels a front end can ask for -- int32 output on the aligned grid, FP64 otherwise, SSA in
// FP64 -- with the default sampler: a third of the instantiations of a built-in model, ~15 s of nvcc.
template <class M>
static int launch_plugin_jump(const JumpLaunch& L0, int method, bool aligned, int dtype, cudaStream_t stream) {
    JumpLaunch L = L0;
    L.lane_shift = method == SODM_JUMP_SSA ? 0 : pick_lane_shift(L.n_traj);
    const unsigned grid = spread_grid(L.n_traj, L.lane_shift);
    if (method == SODM_JUMP_SSA) {
        if (dtype != SODM_F64) return fail(SODM_EUNSUPPORTED, "generated models write SSA output as float64");
        ssa_kernel<M, double><<<grid, JUMP_BLOCK, 0, stream>>>(L);
    } else {
        const size_t smem = tau_leap_smem<M>(L.n_tdpf_values);
        void (*kernel)(JumpLaunch) = nullptr;
        constexpr int T = SODM_WALK_MIN_BLOCKS;
        // SODM_TUNE_LATENCY: the 96-register build (jump_kernels.cuh), offered for the posterior path (int32, aligned grid)
        const bool latency = (L.compat & TUNE_LATENCY) != 0u;
        if (aligned && dtype == SODM_I32) kernel = latency ? tau_leap_kernel<M, int32_t, true, false, LATENCY_MIN_BLOCKS> : tau_leap_kernel<M, int32_t, true, false, T>;
        else if (aligned && dtype == SODM_F64) kernel = tau_leap_kernel<M, double, true, false, T>;
        else if (!aligned && dtype == SODM_F64) kernel = tau_leap_kernel<M, double, false, false, T>;
        else return fail(SODM_EUNSUPPORTED, "generated models write int32 (tau = 2^-k) or float64 output");
        if (smem > 200 * 1024) return fail(SODM_EUNSUPPORTED, "model needs %zu bytes of shared memory per block", smem);
        SODM_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 96 * 1024 ? smem : 96 * 1024)));
        kernel<<<grid, JUMP_BLOCK, smem, stream>>>(L);
    }
    SODM_CUDA_CHECK(cudaGetLastError());
    return SODM_OK;
}

}  // namespace sodm
