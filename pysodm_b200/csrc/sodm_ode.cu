// ODE ensemble kernels and their C-ABI entry point.
//
// Small systems (SIR, SIR-SI, PPBB): one thread per parameter draw, whole state + the seven RK stages in
// registers, FP64, per-trajectory adaptive step control (lanes of a warp run their own step sequences).
// Packed-bed reactor (8*nx coupled equations): 1 / 2 / 4 WARPS per draw, one axial node per lane, the three-point
// stencil through register shuffles (+ a shared-memory edge exchange between the warps of a draw), error norms through
// butterfly reductions -- every lane of a draw takes the same accept/reject decisions, so a trajectory never diverges.
// Replaces /root/reference/src/pySODM/models/base.py:950-957 (the loop over draws in `ODE.sim`) and the
// `solve_ivp` / `_solve_discrete` calls below it (:853-855).
#include "ode_kernels.cuh"

namespace sodm {

// ---- packed-bed reactor: NW warps per draw, ONE axial node (8 equations) per lane ------------------------------------------
// Round 1 gave a lane NPL = ceil(nx / 32) nodes and kept all 9 Runge-Kutta vectors in registers: 72 doubles + 19 per-draw
// coefficients per lane at nx <= 32 (168 registers under a 3-block bound, 9 % of the executed instructions local-memory
// spills, 12 warps per SM), 288 doubles at nx = 100 (everything spilled: 4.6 % of the FP64 peak).  Now
//   * a lane always owns one node; a draw takes NW = 1 / 2 / 4 warps (nx <= 32 / 64 / 128); the warps of a draw exchange their
//     edge nodes and their partial norms through shared memory, one named barrier per exchange (double-buffered), and
//     take identical accept / reject decisions, so a draw never diverges;
//   * the stage derivatives K2 .. K6 live in shared memory ([stage][element][thread]: conflict-free 8-byte columns), the
//     per-draw coefficients too (read by broadcast); registers hold y, the stage input, K1 and K7: 128 registers, 4 blocks.
constexpr int PFR_BLOCK = 128;
#ifndef SODM_PFR_MIN_BLOCKS
#define SODM_PFR_MIN_BLOCKS 4
#endif

struct SharedStages {                      // K2 .. K6 of this lane: element (j, e) at base[(j * 8 + e) * PFR_BLOCK]
    double* base;
    __device__ __forceinline__ double get(int j, int e) const { return base[(j * 8 + e) * PFR_BLOCK]; }
    __device__ __forceinline__ void put(int j, const double* v) {
#pragma unroll
        for (int e = 0; e < 8; ++e) base[(j * 8 + e) * PFR_BLOCK] = v[e];
    }
};

template <int NW>
__device__ __forceinline__ void draw_barrier(int id) {
    if (NW == 1) __syncwarp();
    else if (NW * 32 == PFR_BLOCK) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(NW * 32) : "memory");
}

// Local element e = phase * 4 + species (phase 0 = fluid C_F, 1 = solid C_S) of axial node `node`.
template <int NW>
struct PfrBlockSystem {
    static constexpr int NL = 8;
    const PfrNode::Coef* coef;      // shared, one per draw
    double* edge;                   // shared, this draw: [2 parities][NW][2 sides][4 species]
    double* part;                   // shared, this draw: [2 parities][NW]
    int nx, node, lane, wid, bar;   // wid = warp within the draw; bar = named barrier of the draw
    int px, pr;                     // parities of the two exchange buffers
    __device__ __forceinline__ int n_total() const { return 8 * nx; }
    __device__ __forceinline__ bool active(int) const { return node < nx; }
    __device__ __forceinline__ bool leader() const { return node == 0; }
    __device__ __forceinline__ int global_index(int e) const { return e * nx + node; }
    __device__ __forceinline__ double reduce(double x) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (NW > 1) {
            if (lane == 0) part[pr * NW + wid] = x;
            draw_barrier<NW>(bar);
            x = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) x += part[pr * NW + w];     // same order in every warp: identical decisions
            pr ^= 1;
        }
        return x;
    }
    __device__ __forceinline__ void rhs(double, const double* y, double* dy) {
        double cl[4], cr[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            cl[s] = __shfl_up_sync(0xffffffffu, y[s], 1);
            cr[s] = __shfl_down_sync(0xffffffffu, y[s], 1);
        }
        if (NW > 1) {                                            // the neighbours of a warp's edge lanes live in other warps
            double* mine = edge + ((px * NW + wid) * 2) * 4;
            if (lane == 0 || lane == 31) {
#pragma unroll
                for (int s = 0; s < 4; ++s) mine[(lane == 31 ? 4 : 0) + s] = y[s];
            }
            draw_barrier<NW>(bar);
            if (lane == 0 && wid > 0) {
#pragma unroll
                for (int s = 0; s < 4; ++s) cl[s] = edge[((px * NW + wid - 1) * 2 + 1) * 4 + s];
            }
            if (lane == 31 && wid < NW - 1) {
#pragma unroll
                for (int s = 0; s < 4; ++s) cr[s] = edge[((px * NW + wid + 1) * 2) * 4 + s];
            }
            px ^= 1;
        }
        double dcf[4], dcs[4];
        PfrNode::rhs(*coef, node == 0, node == nx - 1, cl, y, cr, y + 4, dcf, dcs);
        const bool on = node < nx;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            dy[s] = on ? dcf[s] : 0.0;
            dy[4 + s] = on ? dcs[s] : 0.0;
        }
    }
};

// stage vectors a lane parks in shared memory: K2 .. K6 of RK45 (2 of them for RK23), 14 for DOP853 (dop853_scheme.cuh)
__host__ __device__ constexpr int pfr_stage_slots(int method) { return method == SODM_ODE_DOP853 ? SchemeDOP853::NSLOT : 5; }

template <int NW>
constexpr size_t pfr_smem_bytes(int slots) {
    constexpr int draws = PFR_BLOCK / (32 * NW);
    return (size_t)slots * 8 * PFR_BLOCK * sizeof(double) + draws * sizeof(PfrNode::Coef) +
           draws * (2 * NW * 2 * 4 + 2 * NW) * sizeof(double);
}

// RK23 and DOP853 are kernels of their own (GROUP 1, 2; GROUP 0 = RK45 + the fixed-step schemes): inside one common kernel the
// 14 parked stage vectors and 12-term combinations of DOP853 raise the register pressure of every method (RK45: 155 -> 160 ms
// at nx = 25, with spills in the common path).
template <int NW, typename OutT, int GROUP = 0>
__global__ void __launch_bounds__(PFR_BLOCK, SODM_PFR_MIN_BLOCKS) ode_pfr_warp_kernel(const __grid_constant__ OdeLaunch L, int method) {
    extern __shared__ double pfr_smem[];
    constexpr int DRAWS = PFR_BLOCK / (32 * NW);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d = warp / NW;                                     // draw within the block
    const int64_t i = (int64_t)blockIdx.x * DRAWS + d;
    if (i >= L.n_traj) return;                                   // all warps of a draw leave together
    double* stages = pfr_smem;
    PfrNode::Coef* coefs = reinterpret_cast<PfrNode::Coef*>(pfr_smem + pfr_stage_slots(method) * 8 * PFR_BLOCK);
    double* xch = reinterpret_cast<double*>(coefs + DRAWS) + d * (2 * NW * 2 * 4 + 2 * NW);
    PfrBlockSystem<NW> sys;
    sys.coef = coefs + d;
    sys.edge = xch;
    sys.part = xch + 2 * NW * 2 * 4;
    sys.nx = L.nx; sys.lane = lane; sys.wid = warp % NW; sys.node = sys.wid * 32 + lane;
    sys.bar = 1 + d; sys.px = 0; sys.pr = 0;
    if (sys.node == 0) {                                         // the draw's coefficients: formed once, read by broadcast
        double par[PfrNode::NPAR];
        load_ode_params<PfrNode::NPAR>(L, i, par);
        coefs[d] = PfrNode::coefficients(par);
    }
    draw_barrier<NW>(sys.bar);
    double y[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        y[e] = 1.0;                      // lanes without a node: results are never used, but 0 / 0 in their rate law would drag
        if (sys.active(e)) {             // the warp through the FP64 division's slow path at every stage (fdiv.cuh)
            const int s = sys.global_index(e);
            y[e] = L.y0_bcast ? L.y0[s] : L.y0[(int64_t)s * L.ld_in + i];
        }
    }
    if constexpr (GROUP == 2) {
        SharedStages S{stages + threadIdx.x};
        rk45_run<PfrBlockSystem<NW>, OutT, SharedStages, SchemeDOP853>(sys, L, i, y, S);
        return;
    } else if constexpr (GROUP == 1) {
        SharedStages S{stages + threadIdx.x};
        rk45_run<PfrBlockSystem<NW>, OutT, SharedStages, SchemeRK23>(sys, L, i, y, S);
        return;
    }
    if (method == SODM_ODE_RK45) {
        SharedStages S{stages + threadIdx.x};
        rk45_run<PfrBlockSystem<NW>, OutT, SharedStages>(sys, L, i, y, S);
    } else if (method == SODM_ODE_EULER) fixed_step_run<PfrBlockSystem<NW>, OutT, false>(sys, L, i, y);
    else fixed_step_run<PfrBlockSystem<NW>, OutT, true>(sys, L, i, y);
}

template <int NW>
static int launch_pfr(const OdeLaunch& L, int method, int dtype, cudaStream_t stream) {
    constexpr int DRAWS = PFR_BLOCK / (32 * NW);
    const unsigned grid = (unsigned)((L.n_traj + DRAWS - 1) / DRAWS);
    const size_t smem = pfr_smem_bytes<NW>(pfr_stage_slots(method));
    void (*kernel)(OdeLaunch, int);
    if (method == SODM_ODE_DOP853) kernel = dtype == SODM_F64 ? ode_pfr_warp_kernel<NW, double, 2> : ode_pfr_warp_kernel<NW, float, 2>;
    else if (method == SODM_ODE_RK23) kernel = dtype == SODM_F64 ? ode_pfr_warp_kernel<NW, double, 1> : ode_pfr_warp_kernel<NW, float, 1>;
    else kernel = dtype == SODM_F64 ? ode_pfr_warp_kernel<NW, double, 0> : ode_pfr_warp_kernel<NW, float, 0>;
    SODM_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<grid, PFR_BLOCK, smem, stream>>>(L, method);
    SODM_CUDA_CHECK(cudaGetLastError());
    return SODM_OK;
}

}  // namespace sodm

using namespace sodm;

extern "C" int32_t sodm_ode_n_kept(int32_t n_out, int32_t out_stride, int32_t out_last_only) {
    return ode_n_kept(n_out, out_stride, out_last_only);
}

extern "C" int sodm_ode_run(const sodm_ode_args* a) {
    OdeLaunch L;
    sodm_model_info info;
    const int rc = build_ode_launch(a, L, info);
    if (rc != SODM_OK) return rc;
    if (a->n_traj == 0) return SODM_OK;
    cudaStream_t stream = (cudaStream_t)a->stream;
    switch (a->model) {
        case MODEL_SIR_ODE:      return launch_thread<SirOde>(L, a->method, a->out_dtype, stream);
        case MODEL_SIR_RATE_ODE: return launch_thread<SirRateOde>(L, a->method, a->out_dtype, stream);
        case MODEL_SIR_SI_ODE:   return launch_thread<SirSiOde>(L, a->method, a->out_dtype, stream);
        case MODEL_PPBB_ODE:     return launch_thread<PpbbOde>(L, a->method, a->out_dtype, stream);
        case MODEL_PACKED_PFR_ODE:
            if (a->method == SODM_ODE_RADAU || a->method == SODM_ODE_BDF)
                return fail(SODM_EUNSUPPORTED, "Radau / BDF are offered for thread-per-trajectory models (at most %d equations), not for the reactor's node-per-lane layout", IMPLICIT_MAX_N);
            if (L.nx <= 32) return launch_pfr<1>(L, a->method, a->out_dtype, stream);     // warps per draw
            if (L.nx <= 64) return launch_pfr<2>(L, a->method, a->out_dtype, stream);
            return launch_pfr<4>(L, a->method, a->out_dtype, stream);
    }
    return fail(SODM_EMODEL, "sodm_ode_run: model %d not registered", a->model);
}
