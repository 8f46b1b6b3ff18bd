// Jump-process ensemble kernels (tau-leap, SSA) and their launchers, as templates over the model functor: shared by
// the built-in models (sodm_jump.cu) and by generated models compiled into plugin libraries (plugin.cuh).
//
// One thread per trajectory, state and parameters in registers, structure-of-arrays I/O so that a
// warp's loads/stores are coalesced 128-byte transactions, the time-dependent-parameter table (the
// influenza contact matrices) staged once per block in shared memory and read by broadcast.
#pragma once
#include <cstdlib>
#include "jump_setup.cuh"

namespace sodm {

// One warp per block: the hardware then balances a launch over the SMs warp by warp (a strong-scaling shard of 125 000
// trajectories is 0.8 waves: with 128-thread blocks an SM gets 6 or 7 blocks, with 32-thread blocks 26 or 27 warps), and 28
// resident blocks leave 72 registers per thread instead of 64.  Measured (B200, influenza, full season, profiles/r2_shard_ab.txt):
// 125 000 trajectories 16.14 -> 15.53 ms, 10^6 103.19 -> 102.19 ms; 64- and 256-thread blocks are slower than either.
#ifndef SODM_JUMP_BLOCK
#define SODM_JUMP_BLOCK 32
#endif
constexpr int JUMP_BLOCK = SODM_JUMP_BLOCK;
// Resident blocks per SM the compiler must allow for.  The samplers are issue-bound, so resident warps matter more than
// registers (round 1, legacy walk on influenza, 128-thread blocks: 352 / 272 / 262 / 259 ms at 1 / 5 / 6 / 8 blocks).
#ifndef SODM_WALK_MIN_BLOCKS
#define SODM_WALK_MIN_BLOCKS 28
#endif

// shared-memory words of the per-lane draw order (one byte per draw and lane)
template <class M>
__host__ __device__ constexpr int order_words() { return (M::NDRAW * JUMP_BLOCK + 3) / 4; }

// shared-memory words of the queue sampler's item arrays (16 bits per draw and lane)
template <class M>
__host__ __device__ constexpr int queue_words() { return M::NDRAW * JUMP_BLOCK; }

// bytes of dynamic shared memory of a tau-leap block: [NDRAW][JUMP_BLOCK] draw counts | [NDRAW][JUMP_BLOCK] probabilities |
// TDPF table (FP32) | walk sampler: [NDRAW][JUMP_BLOCK] draw order (bytes); queue sampler: [warp][2 * NDRAW * 32] items (16 bits)
template <class M>
static size_t tau_leap_smem(int n_tdpf_values, bool queue = false) {
    return ((size_t)2 * M::NDRAW * JUMP_BLOCK + (size_t)n_tdpf_values * M::TDPF_COUNT + (size_t)(queue ? queue_words<M>() : order_words<M>())) * sizeof(float);
}

// The same kernel is compiled for two register budgets: MINB = 8 resident blocks per SM (64 registers, some spills) for
// launches that fill the GPU, where resident warps buy issue slots, and MINB = 5 (96 registers, no spills), selected by
// SODM_TUNE_LATENCY, for launches of about one wave or less, whose duration is ONE warp's chain of dependent instructions from
// the first step to the last -- there the spills only lengthen that chain (B200, influenza, 168 steps,
// profiles/r2_small_launches.txt: 12 500 trajectories 4.66 -> 3.90 ms, 1024: 3.54 -> 2.92 ms; 10^6: 108.7 -> 112.7 ms).
// It is a FLAG of the caller, not a choice of the launcher: the two builds are the same source, but ptxas contracts
// different multiply-add pairs under the two register budgets, so about 3 in 10^5 trajectories take a different accept /
// reject decision somewhere (measured: 2 of 60 000 final states differ; with -fmad=false none do, at +3.2 % on the headline
// launch).  Same law, different rounding: a run is reproducible and shard-invariant for a given flag, and a launcher that
// switched builds by launch size would make results depend on the number of GPUs.
constexpr int LATENCY_MIN_BLOCKS = 640 / JUMP_BLOCK;      // 20 warps per SM: 96 registers

template <class M, typename OutT, bool ALIGNED, bool QUEUE, int MINB>
__global__ void __launch_bounds__(JUMP_BLOCK, MINB) tau_leap_kernel(const __grid_constant__ JumpLaunch L) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WorkList w;
    w.n = reinterpret_cast<int*>(smem) + warp * 32;
    w.p = smem + M::NDRAW * JUMP_BLOCK + warp * 32;
    w.stride = JUMP_BLOCK;
    w.lane = lane;
    float* table = smem + 2 * M::NDRAW * JUMP_BLOCK;
    const int n_tab = L.n_tdpf_values * M::TDPF_COUNT;
    w.ord = QUEUE ? nullptr : reinterpret_cast<uint8_t*>(table + n_tab) + warp * 32;
    w.queue = QUEUE ? reinterpret_cast<uint16_t*>(table + n_tab) + warp * (2 * M::NDRAW * 32) : nullptr;
    for (int e = threadIdx.x; e < n_tab; e += blockDim.x) table[e] = (float)L.tdpf_table[e];   // ABI speaks FP64, rates are FP32
    __syncthreads();
    // small ensembles leave warp slots idle: spreading the trajectories over more warps (32 >> lane_shift per warp) shortens
    // every warp's sampling loop (fewer lanes to wait for), see pick_lane_shift
    const int lanes_used = 32 >> L.lane_shift;
    const int64_t i = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * lanes_used + lane;
    const bool active = lane < lanes_used && i < L.n_traj;
    w.lanes = __ballot_sync(0xffffffffu, active);
    if (w.lanes == 0u) return;                                    // whole warp past the end
    tau_leap_trajectory<M, OutT, ALIGNED, QUEUE>(L, i, active, table, w);
}

template <class M, typename OutT>
__global__ void __launch_bounds__(JUMP_BLOCK) ssa_kernel(const __grid_constant__ JumpLaunch L) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < L.n_traj) ssa_trajectory<M, OutT>(L, i);
}

// Trajectories per warp for the tau-leap samplers: the largest spread (32, 16, 8 or 4 lanes in use) that keeps the launch below
// a third of the device's resident-warp slots.  Results do not change (Philox streams are keyed by trajectory, not by
// lane).  Measured (profiles/r1_lane_spread_ab.txt, C5 window): a warp's 168 steps take ~6.8 ms however few lanes it has
// (a dependent chain of Philox / log / compare per draw), and above ~10 warps per SM the kernel is issue-bound, where
// every extra warp costs; so spreading only pays for ensembles of a few thousand trajectories (4096: 6.8 -> 5.2 ms).
// SODM_LANE_SHIFT=0..3 in the environment overrides the choice (A/B measurements).
static int pick_lane_shift(int64_t n_traj) {
    static int slots = 0, forced = -2;                       // idempotent initialisation: `forced` is published last
    if (forced == -2) {
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        slots = sms * SODM_WALK_MIN_BLOCKS * (JUMP_BLOCK / 32) / 3;
        const char* e = getenv("SODM_LANE_SHIFT");
        forced = e && *e >= '0' && *e <= '3' ? *e - '0' : -1;
    }
    if (forced >= 0) return forced;
    int shift = 0;
    while (shift < 3 && (n_traj + (32 >> (shift + 1)) - 1) / (32 >> (shift + 1)) <= slots) ++shift;
    return shift;
}

static unsigned spread_grid(int64_t n_traj, int lane_shift) {
    const int64_t warps = (n_traj + (32 >> lane_shift) - 1) / (32 >> lane_shift);
    return (unsigned)((warps + JUMP_BLOCK / 32 - 1) / (JUMP_BLOCK / 32));
}

template <class M, typename OutT>
static int launch_jump(const JumpLaunch& L0, int method, bool aligned, cudaStream_t stream) {
    JumpLaunch L = L0;
    const bool queue = (L.compat & TUNE_QUEUE) != 0u;
    L.lane_shift = method == SODM_JUMP_TAU_LEAP ? pick_lane_shift(L.n_traj) : 0;
    const unsigned grid = spread_grid(L.n_traj, L.lane_shift);
    if (method == SODM_JUMP_TAU_LEAP) {
        const size_t smem = tau_leap_smem<M>(L.n_tdpf_values, queue);
        // the queue sampler exists in the 96-register build only (at 64 registers it spills in its batch loop: 155 vs 118 ms)
        constexpr int T = SODM_WALK_MIN_BLOCKS, Q = LATENCY_MIN_BLOCKS;
        const bool latency = (L.compat & TUNE_LATENCY) != 0u;
        void (*kernel)(JumpLaunch);
        if (queue) kernel = aligned ? tau_leap_kernel<M, OutT, true, true, Q> : tau_leap_kernel<M, OutT, false, true, Q>;
        else if (latency) kernel = aligned ? tau_leap_kernel<M, OutT, true, false, Q> : tau_leap_kernel<M, OutT, false, false, Q>;
        else kernel = aligned ? tau_leap_kernel<M, OutT, true, false, T> : tau_leap_kernel<M, OutT, false, false, T>;
        SODM_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        kernel<<<grid, JUMP_BLOCK, smem, stream>>>(L);
    } else {
        ssa_kernel<M, OutT><<<grid, JUMP_BLOCK, 0, stream>>>(L);
    }
    SODM_CUDA_CHECK(cudaGetLastError());
    return SODM_OK;
}

template <class M>
static int launch_jump_dtype(const JumpLaunch& L, int method, bool aligned, int dtype, cudaStream_t stream) {
    switch (dtype) {
        case SODM_I32: return launch_jump<M, int32_t>(L, method, aligned, stream);
        case SODM_F32: return launch_jump<M, float>(L, method, aligned, stream);
        case SODM_F64: return launch_jump<M, double>(L, method, aligned, stream);
    }
    return fail(SODM_EINVAL, "sodm_jump_run: unknown out_dtype %d", dtype);
}

}  // namespace sodm

