// Thread-per-trajectory ODE ensemble kernel and its launcher, as templates over the model functor: shared by the
// built-in models (sodm_ode.cu) and by generated models compiled into plugin libraries (plugin.cuh).
#pragma once
#include "ode_setup.cuh"

namespace sodm {

constexpr int ODE_BLOCK = 128;

template <class M, typename OutT>
__global__ void __launch_bounds__(ODE_BLOCK) ode_thread_kernel(const __grid_constant__ OdeLaunch L, int method) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < L.n_traj) ode_thread_trajectory<M, OutT>(L, method, i);
}

template <class M>
static int launch_thread(const OdeLaunch& L, int method, int dtype, cudaStream_t stream) {
    const unsigned grid = (unsigned)((L.n_traj + ODE_BLOCK - 1) / ODE_BLOCK);
    if (dtype == SODM_F64) ode_thread_kernel<M, double><<<grid, ODE_BLOCK, 0, stream>>>(L, method);
    else                   ode_thread_kernel<M, float><<<grid, ODE_BLOCK, 0, stream>>>(L, method);
    SODM_CUDA_CHECK(cudaGetLastError());
    return SODM_OK;
}

}  // namespace sodm
