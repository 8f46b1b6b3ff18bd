"""Python-side description of the device functors registered in libsodm.so (flat state / parameter orders).

A pySODM model class is tied to a hand-written `__device__` functor through its `device_model` attribute; this
table says how the class's named states and parameters map onto the functor's flat vectors (states in declaration
order, C-order ravel -- /root/reference/src/pySODM/models/utils.py:45-55 -- and the parameter order documented in
csrc/sodm_core.cu).  `check_against_library()` verifies the sizes against `sodm_model_info_get`.
"""
from collections import OrderedDict

import numpy as np


class DeviceSpec:
    def __init__(self, name, kind, states, params, tdpf_params=(), sized_by=None):
        self.name, self.kind = name, kind
        self.states = OrderedDict(states)          # name -> shape (tuple; may contain the symbol 'nx')
        self.params = OrderedDict(params)          # name -> shape, in FLAT ORDER
        self.tdpf_params = tuple(tdpf_params)      # parameters that may be time dependent (read from the TDPF table)
        self.sized_by = sized_by                   # dimension name whose length is the functor's model_arg

    def resolve(self, model_arg=0):
        def shp(s):
            return tuple(model_arg if d == "nx" else d for d in s)
        return (OrderedDict((k, shp(v)) for k, v in self.states.items()),
                OrderedDict((k, shp(v)) for k, v in self.params.items()))

    def param_offsets(self, model_arg=0):
        _, params = self.resolve(model_arg)
        off, out = 0, OrderedDict()
        for k, s in params.items():
            n = int(np.prod(s)) if s else 1
            out[k] = (off, n, s)
            off += n
        return out, off


_AGE = (4,)
SPECS = {s.name: s for s in [
    DeviceSpec("sir_jump", "jump", [("S", ()), ("I", ()), ("R", ())], [("beta", ()), ("gamma", ())]),
    DeviceSpec("sir_si_jump", "jump",
               [("S", _AGE), ("I", _AGE), ("R", _AGE), ("S_v", ()), ("I_v", ())],
               [("alpha", _AGE), ("beta", ()), ("gamma", ())]),
    DeviceSpec("influenza_1718_jump", "jump",
               [(s, _AGE) for s in ("S", "E", "Ip", "Iud", "Id", "R", "Im_inc")],
               [("alpha", ()), ("beta", ()), ("gamma", ()), ("delta", ()), ("f_ud", _AGE), ("N", (4, 4))],
               tdpf_params=("N",)),
    DeviceSpec("sir_ode", "ode", [("S", ()), ("I", ()), ("R", ())], [("beta", ()), ("gamma", ())]),
    DeviceSpec("sir_rate_ode", "ode", [("S", ()), ("I", ()), ("R", ())], [("beta", ()), ("gamma", ())]),
    DeviceSpec("sir_si_ode", "ode",
               [("S", _AGE), ("I", _AGE), ("R", _AGE), ("S_v", ()), ("I_v", ())],
               [("alpha", _AGE), ("beta", ()), ("gamma", ())]),
    DeviceSpec("ppbb_ode", "ode", [("S", ()), ("A", ()), ("Es", ()), ("W", ())],
               [("c_enzyme", ()), ("Vf_Ks", ()), ("R_AS", ()), ("R_AW", ()), ("R_Es", ()), ("K_eq", ())]),
    DeviceSpec("packed_pfr_ode", "ode", [("C_F", (4, "nx")), ("C_S", (4, "nx"))],
               [("delta_x", ()), ("epsilon", ()), ("u", ()), ("rho_B", ()), ("Vf_Ks", ()), ("R_AS", ()), ("R_AW", ()),
                ("K_eq", ()), ("R_Es", ()), ("kL_a", (4,)), ("D_ax", (4,))], sized_by="x"),
]}


def get(name):
    if name not in SPECS:
        raise NotImplementedError(
            f"no device functor registered under '{name}' (registered: {sorted(SPECS)}).  Leave `device_model` unset to "
            "have the model's own Python functions translated instead (codegen.py); there is no CPU fallback.")
    return SPECS[name]


def check_against_library():
    from . import _ffi
    for name, spec in SPECS.items():
        arg = 25 if spec.sized_by else 0
        info = _ffi.model_info(_ffi.model_id(name), arg)
        states, _ = spec.resolve(arg)
        n_state = sum(int(np.prod(s)) if s else 1 for s in states.values())
        _, n_param = spec.param_offsets(arg)
        assert (info.n_state, info.n_param) == (n_state, n_param), (name, info.n_state, n_state, info.n_param, n_param)
        assert info.kind == (0 if spec.kind == "jump" else 1), name
