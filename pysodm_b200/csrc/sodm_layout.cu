// Output-layout kernel: device layout [row][time][trajectory] -> the reference's user-visible layout
// [trajectory ("draws")][time][state dims] of one state variable
// (/root/reference/src/pySODM/models/base.py:1009-1010 moveaxis, :569-571 concat along "draws").
//
// Pure data movement, HBM-bound: a 32 x 32 shared-memory tile transpose so that both the reads
// (consecutive trajectories) and the writes (consecutive (time, dim) elements of one trajectory) are
// coalesced; dtype conversion is fused.  `dst` may be pinned, device-mapped host memory: the kernel then
// writes the transposed result straight across PCIe, fusing the transpose with the device->host copy.
#include "common.cuh"

namespace sodm {

template <typename S, typename D>
__global__ void __launch_bounds__(256) draws_major_kernel(const S* __restrict__ src, int64_t ld, int r0, int n_dim,
                                                          int n_t, int64_t n_traj, D* __restrict__ dst) {
    __shared__ D tile[32][33];
    const int64_t Q = (int64_t)n_t * n_dim;            // elements per trajectory in dst, q = j*n_dim + d
    const int64_t i0 = (int64_t)blockIdx.x * 32;
    const int64_t q0 = (int64_t)blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;      // 32 x 8
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int64_t q = q0 + ty + 8 * r;
        const int64_t i = i0 + tx;
        if (q < Q && i < n_traj) {
            const int64_t j = q / n_dim, d = q % n_dim;
            tile[ty + 8 * r][tx] = (D)src[((int64_t)(r0 + d) * n_t + j) * ld + i];
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int64_t i = i0 + ty + 8 * r;
        const int64_t q = q0 + tx;
        if (q < Q && i < n_traj) dst[i * Q + q] = tile[tx][ty + 8 * r];
    }
}

template <typename S>
static int launch_layout(const void* src, int64_t ld, int r0, int n_dim, int n_t, int64_t n_traj, void* dst,
                         int dst_dtype, cudaStream_t stream) {
    const int64_t Q = (int64_t)n_t * n_dim;
    dim3 block(32, 8), grid((unsigned)((n_traj + 31) / 32), (unsigned)((Q + 31) / 32));
    if (grid.y > 65535) return fail(SODM_EUNSUPPORTED, "sodm_to_draws_major: n_t * n_dim too large");
    switch (dst_dtype) {
        case SODM_I32: draws_major_kernel<S, int32_t><<<grid, block, 0, stream>>>((const S*)src, ld, r0, n_dim, n_t, n_traj, (int32_t*)dst); break;
        case SODM_F32: draws_major_kernel<S, float><<<grid, block, 0, stream>>>((const S*)src, ld, r0, n_dim, n_t, n_traj, (float*)dst); break;
        case SODM_F64: draws_major_kernel<S, double><<<grid, block, 0, stream>>>((const S*)src, ld, r0, n_dim, n_t, n_traj, (double*)dst); break;
        default: return fail(SODM_EINVAL, "sodm_to_draws_major: unknown dst_dtype %d", dst_dtype);
    }
    SODM_CUDA_CHECK(cudaGetLastError());
    return SODM_OK;
}

// Bitwise OR of a status vector: per-thread OR over a grid-stride loop, warp shuffle, one atomicOr per warp that saw a flag.
__global__ void __launch_bounds__(256) status_or_kernel(const int32_t* __restrict__ status, int64_t n, int32_t* __restrict__ out) {
    int32_t bits = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) bits |= status[i];
    bits = __reduce_or_sync(0xffffffffu, (unsigned)bits);
    if ((threadIdx.x & 31) == 0 && bits) atomicOr(out, bits);
}

}  // namespace sodm

using namespace sodm;

extern "C" int sodm_status_or(const int32_t* status, int64_t n, int32_t* out, void* stream) {
    if (n < 0 || !out || (n > 0 && !status)) return fail(SODM_EINVAL, "sodm_status_or: null buffer or negative length");
    if (n == 0) return SODM_OK;
    const int64_t blocks = (n + 255) / 256;
    status_or_kernel<<<(unsigned)(blocks < 1184 ? blocks : 1184), 256, 0, (cudaStream_t)stream>>>(status, n, out);
    SODM_CUDA_CHECK(cudaGetLastError());
    return SODM_OK;
}

extern "C" int sodm_to_draws_major(const void* src, int32_t src_dtype, int64_t ld, int32_t r0, int32_t n_dim, int32_t n_t,
                                   int64_t n_traj, void* dst, int32_t dst_dtype, void* stream) {
    if (n_traj == 0) return SODM_OK;
    if (!src || !dst) return fail(SODM_EINVAL, "sodm_to_draws_major: null buffer");
    if (n_dim < 1 || n_t < 1 || r0 < 0 || n_traj < 0 || ld < n_traj)
        return fail(SODM_EINVAL, "sodm_to_draws_major: need n_dim, n_t >= 1, r0 >= 0, 0 <= n_traj <= ld");
    if (n_traj == 0) return SODM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    switch (src_dtype) {
        case SODM_I32: return launch_layout<int32_t>(src, ld, r0, n_dim, n_t, n_traj, dst, dst_dtype, s);
        case SODM_F32: return launch_layout<float>(src, ld, r0, n_dim, n_t, n_traj, dst, dst_dtype, s);
        case SODM_F64: return launch_layout<double>(src, ld, r0, n_dim, n_t, n_traj, dst, dst_dtype, s);
    }
    return fail(SODM_EINVAL, "sodm_to_draws_major: unknown src_dtype %d", src_dtype);
}
