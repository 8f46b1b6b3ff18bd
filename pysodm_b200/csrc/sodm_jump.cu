// Jump-process ensemble kernels (tau-leap, SSA) and their C-ABI entry point.
//
// One thread per trajectory, state and parameters in registers, structure-of-arrays I/O so that a
// warp's loads/stores are coalesced 128-byte transactions, the time-dependent-parameter table (the
// influenza contact matrices) staged once per block in shared memory and read by broadcast.
// Replaces /root/reference/src/pySODM/models/base.py:558-566 (the loop over draws in
// `JumpProcess.sim`) and everything it calls (:206-348, :352-440).
#include "jump_kernels.cuh"

using namespace sodm;

extern "C" int64_t sodm_jump_n_steps(double t0, int32_t n_out, double tau) { return jump_n_steps(t0, n_out, tau); }

extern "C" int sodm_jump_run(const sodm_jump_args* a) {
    JumpLaunch L;
    bool aligned = false;
    const int rc0 = build_jump_launch(a, L, aligned);
    if (rc0 != SODM_OK) return rc0;
    if (a->n_traj == 0) return SODM_OK;
    cudaStream_t stream = (cudaStream_t)a->stream;
    int rc;
    switch (a->model) {
        case MODEL_SIR_JUMP:       rc = launch_jump_dtype<SirJump>(L, a->method, aligned, a->out_dtype, stream); break;
        case MODEL_SIR_SI_JUMP:    rc = launch_jump_dtype<SirSiJump>(L, a->method, aligned, a->out_dtype, stream); break;
        case MODEL_INFLUENZA_JUMP: rc = launch_jump_dtype<InfluenzaJump>(L, a->method, aligned, a->out_dtype, stream); break;
        default: rc = fail(SODM_EMODEL, "sodm_jump_run: model %d not registered", a->model);
    }
    return rc;
}
