// Library-wide pieces of the C ABI: error buffer, device check, model registry.
#include <cstring>
#include "common.cuh"
#include "jump_models.cuh"

namespace sodm {

char* error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}

struct ModelEntry {
    const char* name;
    int kind, n_state, n_param, n_channel, tdpf_offset, tdpf_count, supports_ssa;
};

// Flat parameter orders (documented in DESIGN.md):
//   sir_jump / sir_ode / sir_rate_ode : beta, gamma
//   sir_si_jump / sir_si_ode          : alpha[4], beta, gamma
//   influenza_1718_jump               : alpha, beta, gamma, delta, f_ud[4], N[4][4]
//   ppbb_ode                          : c_enzyme, Vf_Ks, R_AS, R_AW, R_Es, K_eq
//   packed_pfr_ode                    : delta_x, epsilon, u, rho_B, Vf_Ks, R_AS, R_AW, K_eq, R_Es, kL_a[4], D_ax[4]
static const ModelEntry kModels[MODEL_COUNT] = {
    {"sir_jump", SODM_KIND_JUMP, SirJump::NS, SirJump::NPAR, SirJump::NCH, 0, 0, 1},
    {"sir_si_jump", SODM_KIND_JUMP, SirSiJump::NS, SirSiJump::NPAR, SirSiJump::NCH, 0, 0, 1},
    {"influenza_1718_jump", SODM_KIND_JUMP, InfluenzaJump::NS, InfluenzaJump::NPAR, InfluenzaJump::NCH,
     InfluenzaJump::TDPF_OFFSET, InfluenzaJump::TDPF_COUNT, 1},
    {"sir_ode", SODM_KIND_ODE, 3, 2, 0, 0, 0, 0},
    {"sir_rate_ode", SODM_KIND_ODE, 3, 2, 0, 0, 0, 0},
    {"sir_si_ode", SODM_KIND_ODE, 14, 6, 0, 0, 0, 0},
    {"ppbb_ode", SODM_KIND_ODE, 4, 6, 0, 0, 0, 0},
    {"packed_pfr_ode", SODM_KIND_ODE, -8 /* 8 * nx */, 17, 0, 0, 0, 0},
};

}  // namespace sodm

using namespace sodm;

extern "C" int sodm_abi_version(void) { return SODM_ABI_VERSION; }

extern "C" const char* sodm_last_error(void) { return error_buffer(); }

extern "C" int sodm_init(int device) {
    int count = 0;
    SODM_CUDA_CHECK(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(SODM_EINVAL, "sodm_init: device %d of %d", device, count);
    cudaDeviceProp prop;
    SODM_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(SODM_EUNSUPPORTED, "sodm_init: device %d is sm_%d%d; this library is built for sm_100a only",
                    device, prop.major, prop.minor);
    SODM_CUDA_CHECK(cudaSetDevice(device));
    return SODM_OK;
}

extern "C" int sodm_peer_enable(int peer_device) {
    int dev = 0, count = 0;
    SODM_CUDA_CHECK(cudaGetDevice(&dev));
    SODM_CUDA_CHECK(cudaGetDeviceCount(&count));
    if (peer_device < 0 || peer_device >= count) return fail(SODM_EINVAL, "sodm_peer_enable: device %d of %d", peer_device, count);
    if (peer_device == dev) return SODM_OK;
    int can = 0;
    SODM_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, dev, peer_device));
    if (!can) return fail(SODM_EUNSUPPORTED, "sodm_peer_enable: device %d cannot address device %d", dev, peer_device);
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return SODM_OK; }
    SODM_CUDA_CHECK(e);
    return SODM_OK;
}

// Peer-visible output buffers.  A CUDA IPC handle names a whole cudaMalloc allocation, so these buffers are the one thing the
// library allocates itself (everything else is caller-owned, include/sodm.h "Conventions").
extern "C" int sodm_peer_alloc(int64_t bytes, void** ptr, unsigned char* handle) {
    if (bytes <= 0 || !ptr || !handle) return fail(SODM_EINVAL, "sodm_peer_alloc: bytes > 0, ptr and handle must be given");
    static_assert(sizeof(cudaIpcMemHandle_t) == SODM_PEER_HANDLE_BYTES, "sodm.h: SODM_PEER_HANDLE_BYTES");
    void* p = nullptr;
    SODM_CUDA_CHECK(cudaMalloc(&p, (size_t)bytes));
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        cudaGetLastError();
        return fail(SODM_ECUDA, "sodm_peer_alloc: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    }
    memcpy(handle, &h, sizeof h);
    *ptr = p;
    return SODM_OK;
}

extern "C" int sodm_peer_open(const unsigned char* handle, void** ptr) {
    if (!handle || !ptr) return fail(SODM_EINVAL, "sodm_peer_open: handle and ptr must be given");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    void* p = nullptr;
    SODM_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr = p;
    return SODM_OK;
}

extern "C" int sodm_peer_close(void* ptr) {
    if (!ptr) return SODM_OK;
    SODM_CUDA_CHECK(cudaIpcCloseMemHandle(ptr));
    return SODM_OK;
}

extern "C" int sodm_peer_free(void* ptr) {
    if (!ptr) return SODM_OK;
    SODM_CUDA_CHECK(cudaFree(ptr));
    return SODM_OK;
}

// The DMA form of the same gather: one strided 2-D copy (copy engines, no SM) of this rank's columns into a peer's buffer.
extern "C" int sodm_peer_copy2d(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t width_bytes,
                                int64_t rows, void* stream) {
    if (!dst || !src || width_bytes <= 0 || rows <= 0 || dst_pitch < width_bytes || src_pitch < width_bytes)
        return fail(SODM_EINVAL, "sodm_peer_copy2d: pointers, 0 < width <= pitches and rows > 0 must be given");
    SODM_CUDA_CHECK(cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src, (size_t)src_pitch, (size_t)width_bytes, (size_t)rows,
                                      cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return SODM_OK;
}

extern "C" int sodm_model_id(const char* name) {
    if (!name) return fail(SODM_EINVAL, "sodm_model_id: null name");
    for (int m = 0; m < MODEL_COUNT; ++m)
        if (std::strcmp(kModels[m].name, name) == 0) return m;
    return fail(SODM_EMODEL, "sodm_model_id: no device functor registered under '%s'", name);
}

extern "C" int sodm_model_info_get(int model, int model_arg, sodm_model_info* out) {
    if (model < 0 || model >= MODEL_COUNT || !out) return fail(SODM_EMODEL, "sodm_model_info_get: bad model %d", model);
    const ModelEntry& e = kModels[model];
    out->kind = e.kind;
    out->n_state = e.n_state;
    if (e.n_state < 0) {
        if (model_arg < 3 || model_arg > 128)
            return fail(SODM_EINVAL, "sodm_model_info_get: %s needs 3 <= nx <= 128 (got %d)", e.name, model_arg);
        out->n_state = -e.n_state * model_arg;
    }
    out->n_param = e.n_param;
    out->n_channel = e.n_channel;
    out->tdpf_offset = e.tdpf_offset;
    out->tdpf_count = e.tdpf_count;
    out->supports_ssa = e.supports_ssa;
    out->reserved = 0;
    return SODM_OK;
}
