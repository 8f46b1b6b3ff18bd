"""ctypes binding of the C ABI in `include/sodm.h` (the only way pysodm_b200 reaches the GPU).

There is NO CPU fallback: if `libsodm.so` is missing or a symbol of the header is not exported, the
import fails loudly.  PyTorch is used by the callers only to own device buffers and streams; no torch
type crosses this boundary -- only raw pointers (`tensor.data_ptr()`) and sizes.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PYSODM_B200_LIB") or os.path.join(_HERE, "libsodm.so")     # override: tuning builds only

# constants mirrored from include/sodm.h
ABI_VERSION = 5
OK, EINVAL, EMODEL, ECUDA, EUNSUPPORTED = 0, -1, -2, -3, -4
KIND_JUMP, KIND_ODE = 0, 1
I32, F32, F64 = 0, 1, 2
ST_BAD_PROB, ST_NONFINITE, ST_STEP_TOO_SMALL, ST_MAX_STEPS, ST_TRUNCATED_Y0 = 1, 2, 4, 8, 16
COMPAT_STAY_QUIRK, COMPAT_SSA_FIRST_MATCH, COMPAT_REFERENCE = 1, 2, 3
TUNE_QUEUE, TUNE_LATENCY = 256, 512
JUMP_TAU_LEAP, JUMP_SSA = 0, 1
ODE_RK45, ODE_EULER, ODE_RK4, ODE_RK23, ODE_DOP853, ODE_RADAU, ODE_BDF = 0, 1, 2, 3, 4, 5, 6
LL_POISSON, LL_NORMAL, LL_LOGNORMAL, LL_NEGBIN = 0, 1, 2, 3
ARG_NONE, ARG_SCALAR, ARG_COLUMN, ARG_FULL = 0, 1, 2, 3
MAX_PARAM, MAX_STATE = 64, 64
MAX_PEERS, PEER_HANDLE_BYTES = 7, 64
NOISE_POISSON, NOISE_GAUSSIAN_ABS, NOISE_GAUSSIAN_REL, NOISE_NEGBIN = 0, 1, 2, 3
NOISE_ST_BAD_MEAN = 1
MAX_RANKS = 4
SELECT_MINMAX, SELECT_HISTOGRAM, SELECT_PICK, SELECT_DECODE, SELECT_BINS = 0, 1, 2, 3, 2048


class ModelInfo(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_state", C.c_int32), ("n_param", C.c_int32), ("n_channel", C.c_int32),
                ("tdpf_offset", C.c_int32), ("tdpf_count", C.c_int32), ("supports_ssa", C.c_int32),
                ("reserved", C.c_int32)]


class JumpArgs(C.Structure):
    _fields_ = [("model", C.c_int32), ("method", C.c_int32), ("compat", C.c_uint32), ("out_dtype", C.c_int32),
                ("n_traj", C.c_int64), ("traj_id0", C.c_int64), ("seed", C.c_uint64),
                ("ld_in", C.c_int64), ("ld_out", C.c_int64),
                ("y0", C.c_void_p), ("y0_broadcast", C.c_int32), ("n_var", C.c_int32),
                ("par", C.c_void_p), ("var_index", C.c_void_p), ("var_values", C.c_void_p),
                ("tdpf_table", C.c_void_p), ("tdpf_index", C.c_void_p),
                ("n_tdpf_values", C.c_int32), ("n_tdpf_index", C.c_int32),
                ("t0", C.c_double), ("n_out", C.c_int32), ("n_sel", C.c_int32), ("tau", C.c_double),
                ("state_sel", C.c_void_p), ("out", C.c_void_p), ("status", C.c_void_p),
                ("max_events", C.c_int64), ("stream", C.c_void_p),
                ("out_peers", C.c_void_p), ("n_out_peers", C.c_int32), ("reserved2", C.c_int32)]


class OdeArgs(C.Structure):
    _fields_ = [("model", C.c_int32), ("model_arg", C.c_int32), ("method", C.c_int32), ("out_dtype", C.c_int32),
                ("n_traj", C.c_int64), ("ld_in", C.c_int64), ("ld_out", C.c_int64),
                ("y0", C.c_void_p), ("y0_broadcast", C.c_int32), ("n_var", C.c_int32),
                ("par", C.c_void_p), ("var_index", C.c_void_p), ("var_values", C.c_void_p),
                ("t0", C.c_double), ("n_out", C.c_int32), ("out_stride", C.c_int32),
                ("out_last_only", C.c_int32), ("reserved", C.c_int32),
                ("rtol", C.c_double), ("atol", C.c_double), ("tau", C.c_double),
                ("out", C.c_void_p), ("nfev", C.c_void_p), ("status", C.c_void_p),
                ("max_steps", C.c_int64), ("stream", C.c_void_p)]


class LoglikArgs(C.Structure):
    _fields_ = [("kind", C.c_int32), ("traj_dtype", C.c_int32), ("n_traj", C.c_int64), ("ld", C.c_int64),
                ("traj", C.c_void_p), ("n_out", C.c_int32), ("n_col", C.c_int32),
                ("col_ptr", C.c_void_p), ("col_rows", C.c_void_p), ("col_weights", C.c_void_p),
                ("t0", C.c_double), ("n_t", C.c_int32), ("arg_mode", C.c_int32),
                ("t_data", C.c_void_p), ("ydata", C.c_void_p), ("ll_arg", C.c_void_p),
                ("weight", C.c_double), ("lp", C.c_void_p), ("status", C.c_void_p), ("stream", C.c_void_p)]


class NoiseArgs(C.Structure):
    _fields_ = [("kind", C.c_int32), ("param", C.c_double), ("in_", C.c_void_p), ("in_dtype", C.c_int32),
                ("out", C.c_void_p), ("out_dtype", C.c_int32), ("n_cells", C.c_int64), ("n_traj", C.c_int64),
                ("ld", C.c_int64), ("traj_id0", C.c_int64), ("seed", C.c_uint64), ("status", C.c_void_p),
                ("stream", C.c_void_p)]


class SelectState(C.Structure):
    _fields_ = [("lo", C.c_void_p), ("hi", C.c_void_p), ("nan_flag", C.c_void_p), ("settled", C.c_void_p),
                ("prefix", C.c_void_p), ("remaining", C.c_void_p)]


# every symbol include/sodm.h declares: name -> (restype, argtypes)
class PsoArgs(C.Structure):
    _fields_ = [("n_dim", C.c_int32), ("n", C.c_int64), ("ld", C.c_int64), ("x", C.c_void_p), ("v", C.c_void_p),
                ("p", C.c_void_p), ("g", C.c_void_p), ("lb", C.c_void_p), ("ub", C.c_void_p), ("omega", C.c_double),
                ("phip", C.c_double), ("phig", C.c_double), ("seed", C.c_uint64), ("iteration", C.c_int64),
                ("member_id0", C.c_int64), ("stream", C.c_void_p)]


class WalkerMoveArgs(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_dim", C.c_int32), ("n_walkers", C.c_int64), ("ld", C.c_int64), ("x", C.c_void_p),
                ("mine", C.c_void_p), ("n_mine", C.c_int64), ("others", C.c_void_p), ("n_others", C.c_int64),
                ("param", C.c_double), ("sigma", C.c_double), ("q", C.c_void_p), ("ld_q", C.c_int64),
                ("log_factor", C.c_void_p), ("chol", C.c_void_p), ("white", C.c_void_p), ("ld_white", C.c_int64), ("seed", C.c_uint64), ("iteration", C.c_int64), ("stream", C.c_void_p)]


POPULATION_MAX_DIM = 16
MOVE_STRETCH, MOVE_DE, MOVE_SNOOKER, MOVE_KDE = 0, 1, 2, 3
WALKERS_ST_NAN = 1

SYMBOLS = {
    "sodm_abi_version": (C.c_int, []),
    "sodm_last_error": (C.c_char_p, []),
    "sodm_init": (C.c_int, [C.c_int]),
    "sodm_model_id": (C.c_int, [C.c_char_p]),
    "sodm_model_info_get": (C.c_int, [C.c_int, C.c_int, C.POINTER(ModelInfo)]),
    "sodm_jump_run": (C.c_int, [C.POINTER(JumpArgs)]),
    "sodm_peer_enable": (C.c_int, [C.c_int]),
    "sodm_peer_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p), C.c_char_p]),
    "sodm_peer_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "sodm_peer_close": (C.c_int, [C.c_void_p]),
    "sodm_peer_free": (C.c_int, [C.c_void_p]),
    "sodm_peer_copy2d": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]),
    "sodm_jump_n_steps": (C.c_int64, [C.c_double, C.c_int32, C.c_double]),
    "sodm_status_or": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "sodm_ode_run": (C.c_int, [C.POINTER(OdeArgs)]),
    "sodm_ode_n_kept": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32]),
    "sodm_loglik": (C.c_int, [C.POINTER(LoglikArgs)]),
    "sodm_to_draws_major": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                      C.c_void_p, C.c_int32, C.c_void_p]),
    "sodm_noise": (C.c_int, [C.POINTER(NoiseArgs)]),
    "sodm_draws_sum": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "sodm_draws_order_stats": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int64,
                                         C.POINTER(C.c_int64), C.c_int32, C.c_void_p, C.c_void_p]),
    "sodm_pso_move": (C.c_int, [C.POINTER(PsoArgs)]),
    "sodm_pso_best": (C.c_int, [C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sodm_kde_logpdf": (C.c_int, [C.c_int32, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                                  C.c_double, C.c_void_p, C.c_void_p]),
    "sodm_walkers_propose": (C.c_int, [C.POINTER(WalkerMoveArgs)]),
    "sodm_population_clip": (C.c_int, [C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "sodm_walkers_accept": (C.c_int, [C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "sodm_draws_select_phase": (C.c_int, [C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                          C.POINTER(SelectState), C.c_void_p, C.c_void_p, C.c_void_p]),
}


class SodmError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libsodm error {code}: {message}")
        self.code = code


_lib = None


def load(path=None):
    """Load libsodm.so (once) and bind every symbol of the header.  Raises if anything is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: the CUDA library has not been built. Run `python -m pysodm_b200.build` "
            "(or `__graft_entry__.build()`). pysodm_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    for name, (restype, argtypes) in SYMBOLS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise ImportError(f"{path} does not export `{name}` declared in include/sodm.h") from exc
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.sodm_abi_version() != ABI_VERSION:
        raise ImportError(f"{path}: ABI version {lib.sodm_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc, lib=None):
    if rc < 0:
        lib = lib or load()
        raise SodmError(rc, lib.sodm_last_error().decode("utf-8", "replace"))
    return rc


def model_id(name: str) -> int:
    lib = load()
    return check(lib.sodm_model_id(name.encode()), lib)


def model_info(model: int, model_arg: int = 0) -> ModelInfo:
    lib = load()
    info = ModelInfo()
    check(lib.sodm_model_info_get(model, model_arg, C.byref(info)), lib)
    return info
