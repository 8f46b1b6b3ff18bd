// Thread-per-trajectory ODE ensemble kernel and its launcher, as templates over the model functor: shared by the
// built-in models (sodm_ode.cu) and by generated models compiled into plugin libraries (plugin.cuh).
#pragma once
#include <type_traits>
#include "ode_setup.cuh"

namespace sodm {

constexpr int ODE_BLOCK = 128;

// RK45 output windows (ode_run.cuh, rk45_run_windowed): systems of at most ODE_WINDOW_MAX_N equations park W grid points per
// thread in shared memory and store them warp-wide; W = as many as fit in ODE_WINDOW_BYTES per block, at most 16 (a lane that
// runs ahead of its warp waits at most one window, so a longer window buys nothing).  Larger systems keep the direct stores:
// with one or two points per window the lanes would step in lockstep per grid point.
constexpr int ODE_WINDOW_MAX_N = 8;
constexpr int ODE_WINDOW_BYTES = 64 * 1024;
template <class M>
__host__ __device__ constexpr int ode_window() {
    if (M::N > ODE_WINDOW_MAX_N) return 0;
    const int w = ODE_WINDOW_BYTES / (M::N * ODE_BLOCK * (int)sizeof(double));
    return w > 16 ? 16 : w;
}

template <class M, typename OutT>
__global__ void __launch_bounds__(ODE_BLOCK) ode_thread_kernel(const __grid_constant__ OdeLaunch L, int method) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < L.n_traj) ode_thread_trajectory<M, OutT, ode_window<M>() == 0>(L, method, i);
}

// the same for RK45 with windowed output: every thread of the block takes part in the warp-wide stores, owner of a trajectory or not
#ifndef SODM_ODE_WINDOWED_MIN_BLOCKS
#define SODM_ODE_WINDOWED_MIN_BLOCKS 4      // 128 registers; measured on C1 at 4 x 10^6 draws: 11.5 ms unbounded (162 registers), 10.0 ms at 4, 13.4 ms at 5
#endif
// a scheme may ask for its own residency: the implicit steppers carry a Jacobian and its factorisations per thread and run
// faster with 2 blocks x 255 registers than with 4 x 128 (profiles/r2_implicit_minb.txt: Radau 5.16 -> 4.25 ms, BDF 11.9 -> 10.3 ms)
template <class Scheme, class = void> struct windowed_min_blocks { static constexpr int value = SODM_ODE_WINDOWED_MIN_BLOCKS; };
template <class Scheme> struct windowed_min_blocks<Scheme, std::void_t<decltype(Scheme::WINDOWED_MIN_BLOCKS)>> {
    static constexpr int value = Scheme::WINDOWED_MIN_BLOCKS;
};
template <class M, typename OutT, class Scheme = SchemeRK45>
__global__ void __launch_bounds__(ODE_BLOCK, windowed_min_blocks<Scheme>::value) ode_thread_rk45_windowed_kernel(const __grid_constant__ OdeLaunch L) {
    extern __shared__ double ode_window_smem[];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    ode_thread_trajectory_windowed<M, OutT, Scheme>(L, i, i < L.n_traj, ode_window_smem + threadIdx.x, ode_window<M>(), ODE_BLOCK);
}

// Radau / BDF of a system without output windows (implicit.cuh): direct stores, one kernel per method
template <class M, typename OutT, class Scheme>
__global__ void __launch_bounds__(ODE_BLOCK) ode_thread_implicit_kernel(const __grid_constant__ OdeLaunch L) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < L.n_traj) ode_thread_trajectory_scheme<M, OutT, Scheme>(L, i);
}

template <class M, typename OutT>
static int launch_thread_implicit(const OdeLaunch& L, int method, cudaStream_t stream) {
    if constexpr (M::N > IMPLICIT_MAX_N) {
        return fail(SODM_EUNSUPPORTED, "Radau / BDF keep a dense Jacobian per thread: at most %d equations (model has %d)", IMPLICIT_MAX_N, M::N);
    } else {
        const unsigned grid = (unsigned)((L.n_traj + ODE_BLOCK - 1) / ODE_BLOCK);
        if constexpr (ode_window<M>() > 0) {
            constexpr int smem = ode_window<M>() * M::N * ODE_BLOCK * (int)sizeof(double);
            void (*kernel)(OdeLaunch) = method == SODM_ODE_RADAU ? ode_thread_rk45_windowed_kernel<M, OutT, SchemeRadau>
                                                                 : ode_thread_rk45_windowed_kernel<M, OutT, SchemeBDF>;
            SODM_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            kernel<<<grid, ODE_BLOCK, smem, stream>>>(L);
        } else {
            void (*kernel)(OdeLaunch) = method == SODM_ODE_RADAU ? ode_thread_implicit_kernel<M, OutT, SchemeRadau>
                                                                 : ode_thread_implicit_kernel<M, OutT, SchemeBDF>;
            kernel<<<grid, ODE_BLOCK, 0, stream>>>(L);
        }
        SODM_CUDA_CHECK(cudaGetLastError());
        return SODM_OK;
    }
}

template <class M, typename OutT>
static int launch_thread_dtype(const OdeLaunch& L, int method, cudaStream_t stream) {
    const unsigned grid = (unsigned)((L.n_traj + ODE_BLOCK - 1) / ODE_BLOCK);
    if (method == SODM_ODE_RADAU || method == SODM_ODE_BDF) return launch_thread_implicit<M, OutT>(L, method, stream);
    if constexpr (ode_window<M>() > 0) {
        if (method == SODM_ODE_RK45 || method == SODM_ODE_RK23 || method == SODM_ODE_DOP853) {
            constexpr int smem = ode_window<M>() * M::N * ODE_BLOCK * (int)sizeof(double);
            void (*kernel)(OdeLaunch) = method == SODM_ODE_RK45 ? ode_thread_rk45_windowed_kernel<M, OutT, SchemeRK45>
                                      : method == SODM_ODE_RK23 ? ode_thread_rk45_windowed_kernel<M, OutT, SchemeRK23>
                                                                : ode_thread_rk45_windowed_kernel<M, OutT, SchemeDOP853>;
            SODM_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            kernel<<<grid, ODE_BLOCK, smem, stream>>>(L);
            SODM_CUDA_CHECK(cudaGetLastError());
            return SODM_OK;
        }
    }
    ode_thread_kernel<M, OutT><<<grid, ODE_BLOCK, 0, stream>>>(L, method);
    SODM_CUDA_CHECK(cudaGetLastError());
    return SODM_OK;
}

template <class M>
static int launch_thread(const OdeLaunch& L, int method, int dtype, cudaStream_t stream) {
    if (dtype == SODM_F64) return launch_thread_dtype<M, double>(L, method, stream);
    return launch_thread_dtype<M, float>(L, method, stream);
}

}  // namespace sodm
