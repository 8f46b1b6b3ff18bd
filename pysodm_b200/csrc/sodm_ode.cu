// ODE ensemble kernels and their C-ABI entry point.
//
// Small systems (SIR, SIR-SI, PPBB): one thread per parameter draw, whole state + the seven RK stages in
// registers, FP64, per-trajectory adaptive step control (lanes of a warp run their own step sequences).
// Packed-bed reactor (8*nx coupled equations): one WARP per draw, NPL consecutive axial nodes per lane, the
// three-point stencil through register shuffles, error norms through butterfly reductions -- every lane
// of the warp takes the same accept/reject decisions, so a trajectory never diverges.
// Replaces /root/reference/src/pySODM/models/base.py:950-957 (the loop over draws in `ODE.sim`) and the
// `solve_ivp` / `_solve_discrete` calls below it (:853-855).
#include "ode_kernels.cuh"

namespace sodm {

// resident blocks per SM for the reactor kernel: 248 registers / 2 blocks without the bound; measured on B200 (1e5 draws,
// nx = 25): 330 / 294 / 310 ms at 2 / 3 / 4 blocks -- the kernel waits on FP64 dependency chains, so a third block helps
// more than the spills cost
#ifndef SODM_PFR_MIN_BLOCKS
#define SODM_PFR_MIN_BLOCKS 3
#endif

// One warp per trajectory; lane l owns axial nodes l*NPL .. l*NPL+NPL-1.
// Local element e = (phase*4 + species)*NPL + k  (phase 0 = fluid C_F, 1 = solid C_S; k = node within lane).
template <int NPL>
struct PfrWarpSystem {
    static constexpr int NL = 8 * NPL;
    PfrNode::Coef coef;
    int nx, lane;
    __device__ __forceinline__ int n_total() const { return 8 * nx; }
    __device__ __forceinline__ bool active(int e) const { return lane * NPL + (e % NPL) < nx; }
    __device__ __forceinline__ bool leader() const { return lane == 0; }
    __device__ __forceinline__ int global_index(int e) const { return (e / NPL) * nx + lane * NPL + (e % NPL); }
    __device__ __forceinline__ double reduce(double x) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        return x;
    }
    __device__ __forceinline__ void rhs(double, const double* y, double* dy) const {
        double from_left[4], from_right[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            from_left[s] = __shfl_up_sync(0xffffffffu, y[s * NPL + NPL - 1], 1);    // last node of lane-1
            from_right[s] = __shfl_down_sync(0xffffffffu, y[s * NPL], 1);           // first node of lane+1
        }
#pragma unroll
        for (int k = 0; k < NPL; ++k) {
            const int j = lane * NPL + k;
            double cl[4], cm[4], cr[4], cs[4], dcf[4], dcs[4];
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                cm[s] = y[s * NPL + k];
                cl[s] = k > 0 ? y[s * NPL + k - 1] : from_left[s];
                cr[s] = k < NPL - 1 ? y[s * NPL + k + 1] : from_right[s];
                cs[s] = y[(4 + s) * NPL + k];
            }
            PfrNode::rhs(coef, j == 0, j == nx - 1, cl, cm, cr, cs, dcf, dcs);
            const bool on = j < nx;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                dy[s * NPL + k] = on ? dcf[s] : 0.0;
                dy[(4 + s) * NPL + k] = on ? dcs[s] : 0.0;
            }
        }
    }
};

template <int NPL, typename OutT>
__global__ void __launch_bounds__(ODE_BLOCK, SODM_PFR_MIN_BLOCKS) ode_pfr_warp_kernel(const __grid_constant__ OdeLaunch L, int method) {
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= L.n_traj) return;                                   // whole warps leave together
    double par[PfrNode::NPAR];
    load_ode_params<PfrNode::NPAR>(L, i, par);
    PfrWarpSystem<NPL> sys{PfrNode::coefficients(par), L.nx, lane};
    double y[8 * NPL];
#pragma unroll
    for (int e = 0; e < 8 * NPL; ++e) {
        y[e] = 0.0;
        if (sys.active(e)) {
            const int s = sys.global_index(e);
            y[e] = L.y0_bcast ? L.y0[s] : L.y0[(int64_t)s * L.ld_in + i];
        }
    }
    if (method == SODM_ODE_RK45) rk45_run<PfrWarpSystem<NPL>, OutT>(sys, L, i, y);
    else if (method == SODM_ODE_EULER) fixed_step_run<PfrWarpSystem<NPL>, OutT, false>(sys, L, i, y);
    else fixed_step_run<PfrWarpSystem<NPL>, OutT, true>(sys, L, i, y);
}

template <int NPL>
static int launch_pfr(const OdeLaunch& L, int method, int dtype, cudaStream_t stream) {
    const int64_t threads = L.n_traj * 32;
    const unsigned grid = (unsigned)((threads + ODE_BLOCK - 1) / ODE_BLOCK);
    if (dtype == SODM_F64) ode_pfr_warp_kernel<NPL, double><<<grid, ODE_BLOCK, 0, stream>>>(L, method);
    else                   ode_pfr_warp_kernel<NPL, float><<<grid, ODE_BLOCK, 0, stream>>>(L, method);
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
            if (L.nx <= 32) return launch_pfr<1>(L, a->method, a->out_dtype, stream);
            if (L.nx <= 64) return launch_pfr<2>(L, a->method, a->out_dtype, stream);
            return launch_pfr<4>(L, a->method, a->out_dtype, stream);
    }
    return fail(SODM_EMODEL, "sodm_ode_run: model %d not registered", a->model);
}
