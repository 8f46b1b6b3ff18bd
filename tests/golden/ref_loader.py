"""Import the UNMODIFIED reference (twallema/pySODM at /root/reference) for fixture generation.

Harness glue only -- no reference arithmetic is restated here. The reference imports `xarray`
at module top (`src/pySODM/models/base.py:3`, `optimization/objective_functions.py:5`) and
xarray is not installed in this image, so a three-attribute stub module is placed in
`sys.modules` and `_output_to_xarray_dataset` (`base.py:969`) is replaced by the identity so
that `model.sim()` returns the raw `{'y': (n_flat, n_t), 't': ...}` dict (or a list of them for
N > 1, through the stubbed `xarray.concat`, `base.py:569-571`).

This only works in the build container (where /root/reference is mounted); nothing in the
`-m gpu` tests, `smoke()` or `bench.py` may import this module.
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PYSODM_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "pySODM"))


def _concat(pair, dim):
    head, tail = pair
    if not isinstance(head, list):
        head = [head]
    return head + [tail]


def load():
    """Returns (base, objective_functions) modules of the unmodified reference."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "xarray" not in sys.modules:
        try:
            import xarray  # noqa: F401  (use the real thing when present)
        except ImportError:
            stub = types.ModuleType("xarray")
            stub.Dataset = object
            stub.DataArray = object
            stub.concat = _concat
            stub.__pysodm_stub__ = True
            sys.modules["xarray"] = stub
    src = os.path.join(REFERENCE_ROOT, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    import pySODM.models.base as base
    import pySODM.optimization.objective_functions as objf
    if getattr(sys.modules["xarray"], "__pysodm_stub__", False):
        base._output_to_xarray_dataset = lambda output, *a, **k: output
    return base, objf


def load_tutorial_module(relpath: str, name: str):
    """Import a tutorial file (e.g. 'tutorials/influenza_1718/models.py') by path."""
    load()
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod          # so that multiprocessing can pickle the model classes
    spec.loader.exec_module(mod)
    return mod
