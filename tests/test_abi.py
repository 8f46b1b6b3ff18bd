"""The C-ABI library loads and exports every symbol include/sodm.h declares (no compute calls: CPU only)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from pysodm_b200 import _ffi, build
    build.build_library()
    return _ffi.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "sodm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sodm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    from pysodm_b200 import _ffi
    declared = _declared_symbols()
    assert set(declared) == set(_ffi.SYMBOLS), (declared, sorted(_ffi.SYMBOLS))
    raw = C.CDLL(_ffi.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"libsodm.so does not export {name}"
    assert lib.sodm_abi_version() == _ffi.ABI_VERSION


def test_struct_layout_matches_header():
    """ctypes mirrors must have the C struct sizes (compiled here with gcc from the header itself)."""
    import subprocess
    import tempfile
    from pysodm_b200 import _ffi
    src = ('#include <stdio.h>\n#include "sodm.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(sodm_model_info),'
           ' sizeof(sodm_jump_args), sizeof(sodm_ode_args), sizeof(sodm_loglik_args), sizeof(sodm_noise_args), sizeof(sodm_select_state)); return 0;}\n')
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(_ffi.ModelInfo), C.sizeof(_ffi.JumpArgs), C.sizeof(_ffi.OdeArgs),
                     C.sizeof(_ffi.LoglikArgs), C.sizeof(_ffi.NoiseArgs), C.sizeof(_ffi.SelectState)]


def test_constants_match_header():
    """Every SODM_* constant of include/sodm.h that the ctypes binding mirrors has the header's value (printed by a C program
    compiled from the header): method ids (incl. the implicit ODE methods), status bits, dtypes, error codes, limits."""
    import re
    import subprocess
    import tempfile
    from pysodm_b200 import _ffi
    header = open(os.path.join(ROOT, "include", "sodm.h")).read()
    names = [n for n in re.findall(r"^#define (SODM_[A-Z0-9_]+)[ \t]+[-0-9(]", header, flags=re.M) if hasattr(_ffi, n[5:])]
    assert {"SODM_ODE_RK45", "SODM_ODE_RADAU", "SODM_ODE_BDF", "SODM_ST_MAX_STEPS", "SODM_ABI_VERSION", "SODM_F64"} <= set(names)
    src = '#include <stdio.h>\n#include "sodm.h"\nint main(void){' + "".join(f'printf("%ld\\n", (long)({n}));' for n in names) + "return 0;}\n"
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "k.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "k")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        values = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert len(values) == len(names)
    for n, v in zip(names, values):
        assert getattr(_ffi, n[5:]) == v, (n, v, getattr(_ffi, n[5:]))


def test_model_registry(lib):
    from pysodm_b200 import _ffi
    want = {"sir_jump": (0, 3, 2), "sir_si_jump": (0, 14, 6), "influenza_1718_jump": (0, 28, 24),
            "sir_ode": (1, 3, 2), "sir_rate_ode": (1, 3, 2), "sir_si_ode": (1, 14, 6), "ppbb_ode": (1, 4, 6)}
    for name, (kind, ns, npar) in want.items():
        info = _ffi.model_info(_ffi.model_id(name))
        assert (info.kind, info.n_state, info.n_param) == (kind, ns, npar), name
    flu = _ffi.model_info(_ffi.model_id("influenza_1718_jump"))
    assert (flu.tdpf_offset, flu.tdpf_count, flu.n_channel) == (8, 16, 24)
    pfr = _ffi.model_info(_ffi.model_id("packed_pfr_ode"), 25)
    assert (pfr.n_state, pfr.n_param) == (200, 17)
    with pytest.raises(_ffi.SodmError, match="no device functor"):
        _ffi.model_id("no_such_model")
    with pytest.raises(_ffi.SodmError, match="nx"):
        _ffi.model_info(_ffi.model_id("packed_pfr_ode"), 1000)


def test_argument_validation_without_gpu(lib):
    """Bad requests are rejected before any CUDA call."""
    from pysodm_b200 import _ffi
    a = _ffi.JumpArgs()
    a.model = _ffi.model_id("sir_ode")
    assert lib.sodm_jump_run(C.byref(a)) == _ffi.EMODEL
    a.model = _ffi.model_id("sir_jump")
    a.n_traj, a.ld_in, a.ld_out, a.n_out = 4, 4, 4, 0
    assert lib.sodm_jump_run(C.byref(a)) == _ffi.EINVAL
    assert b"n_out" in lib.sodm_last_error()
    assert lib.sodm_jump_n_steps(0.0, 162, 0.5) == 322
    n, t = 0, 0.0
    while t < 45:
        t, n = t + 0.3, n + 1
    assert lib.sodm_jump_n_steps(0.0, 46, 0.3) == n            # FP64 accumulation (Q5), not round(45 / 0.3)
    assert lib.sodm_ode_n_kept(8001, 100, 0) == 81 and lib.sodm_ode_n_kept(8001, 1, 1) == 1
    o = _ffi.OdeArgs()
    o.model = _ffi.model_id("sir_jump")
    assert lib.sodm_ode_run(C.byref(o)) == _ffi.EMODEL
    ll = _ffi.LoglikArgs()
    ll.kind = 9
    assert lib.sodm_loglik(C.byref(ll)) == _ffi.EINVAL


def test_missing_library_fails_loudly(tmp_path):
    from pysodm_b200 import _ffi
    with pytest.raises(ImportError, match="no CPU fallback"):
        _ffi.load(str(tmp_path / "libsodm.so"))
