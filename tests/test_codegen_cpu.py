"""Generic models, CPU side: the tracer (codegen.py), the generated translation units compile for sm_100a, the plugin
libraries export include/sodm_plugin.h, and the generated functors -- evaluated on the host -- reproduce the Python
functions and the oracle's tau-leap law.  No GPU involved; nvcc is."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gen_models as gm
from oracle import jump as ojump
from pysodm_b200 import codegen, plugin

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tracer_records_numpy_arithmetic():
    g = codegen.Graph()
    x = codegen._sym_array(g, "y", (3,), "f", 0)
    M = np.arange(9.0).reshape(3, 3)
    out = np.exp(-(M @ x)) / np.sum(x) + x ** 2 * np.array([1.0, 0.0, 2.0])
    lines, exprs = codegen.emit(g, list(out), "double", lambda n: f"{n[0]}[{n[1]}]")
    src = "\n".join(lines)
    assert "exp(" in src and len(exprs) == 3
    # the same expressions, evaluated through a tiny interpreter over the DAG, equal NumPy's on numbers
    vals = np.array([0.3, 1.7, 2.2])

    def ev(i, memo={}):
        op, args, _ = g.nodes[i]
        if op == "in":
            return vals[args[0][1]]
        if op == "const":
            return float(args[0])
        a = [ev(j) for j in args]
        return {"add": lambda: a[0] + a[1], "sub": lambda: a[0] - a[1], "mul": lambda: a[0] * a[1], "div": lambda: a[0] / a[1],
                "neg": lambda: -a[0], "exp": lambda: np.exp(a[0]), "pow": lambda: a[0] ** a[1]}[op]()
    want = np.exp(-(M @ vals)) / np.sum(vals) + vals ** 2 * np.array([1.0, 0.0, 2.0])
    assert np.allclose([ev(o.i) for o in out], want, rtol=1e-15)
    # hash-consing: the shared denominator is one node
    assert sum(1 for n in g.nodes if n[0] == "div") == 3


def test_tracer_refuses_what_it_cannot_translate():
    shapes, pars = {"S": (), "I": ()}, {"beta": ()}

    def needs_a_number(t, S, I, beta):
        return -beta * S * np.arange(3.0)[int(I)], beta * I
    with pytest.raises(codegen.TraceError, match="cannot become a Python number"):
        codegen.trace_ode(needs_a_number, shapes, pars)

    def calendar(t, S, I, beta):
        return (-beta * S if t.weekday() < 5 else 0.0), beta * I
    with pytest.raises(codegen.TraceError, match="calendar logic"):
        codegen.trace_ode(calendar, shapes, {"beta": (), "__t0__": ()}, date_unit_us=86400e6)

    def wrong_shape(t, S, I, beta):
        return np.ones(3) * S, I
    with pytest.raises(codegen.TraceError, match="shape"):
        codegen.trace_ode(wrong_shape, shapes, pars)

    def rates(t, S, I, beta):
        return {"S": [beta * I]}

    def apply_uses_time(t, tau, T, S, I, beta):
        return S - T["S"][0] * (t < 3), I + T["S"][0]
    with pytest.raises(codegen.TraceError):
        codegen.jump_source(rates, apply_uses_time, shapes, pars)

    def rates_ok(t, S, I, beta):
        return {"S": [beta * I]}

    def apply_frac(t, tau, T, S, I, beta):
        return S - 0.5 * T["S"][0], I + T["S"][0]
    with pytest.raises(codegen.TraceError, match="integral"):
        codegen.jump_source(rates_ok, apply_frac, shapes, pars)
    with pytest.raises(NotImplementedError):                      # still no CPU fallback: a model must translate or fail
        class Bad(gm.ODE):
            states, parameters = ["S", "I"], ["beta"]
            integrate = staticmethod(needs_a_number)
        Bad({"S": 1.0, "I": 1.0}, {"beta": 0.1})


def test_branches_on_model_values_become_selections():
    def f(t, S, I, beta, k):
        if t < 0:
            b = beta
        elif t < 10 and S > 5:
            b = beta / 1.5
        else:
            b = beta * k
        inf = np.where(I > 1, b * S * I, 0.0)
        return -inf, inf
    g, outs = codegen.trace_ode(f, {"S": (), "I": ()}, {"beta": (), "k": ()})
    lines, exprs = codegen.emit(g, outs, "double", lambda n: "t" if n[0] == "t" else f"{n[0]}[{n[1]}]")
    assert sum(" ? " in ln for ln in lines) >= 6 and any("t < 10.0" in ln for ln in lines)

    # evaluate the DAG on numbers: every path of the Python function is reproduced
    def ev(i, env, memo):
        if i in memo:
            return memo[i]
        op, args, _ = g.nodes[i]
        if op == "in":
            r = env[args[0]]
        elif op == "const":
            r = args[0]
        else:
            a = [ev(j, env, memo) for j in args]
            r = {"add": lambda: a[0] + a[1], "sub": lambda: a[0] - a[1], "mul": lambda: a[0] * a[1], "div": lambda: a[0] / a[1],
                 "neg": lambda: -a[0], "lt": lambda: a[0] < a[1], "gt": lambda: a[0] > a[1], "le": lambda: a[0] <= a[1],
                 "ge": lambda: a[0] >= a[1], "select": lambda: a[1] if a[0] else a[2]}[op]()
        memo[i] = r
        return r
    for t, S, I in [(-1.0, 9.0, 3.0), (5.0, 9.0, 3.0), (5.0, 2.0, 3.0), (50.0, 9.0, 3.0), (5.0, 9.0, 0.5)]:
        env = {("t", 0): t, ("y", 0): S, ("y", 1): I, ("par", 0): 0.3, ("par", 1): 2.0}
        got = [ev(o.i, env, {}) for o in outs]
        assert np.allclose(got, f(t, S, I, 0.3, 2.0), rtol=1e-15)


def test_ode_time_dependent_parameter_functions_translate_and_match_the_reference(golden):
    """The translated composite (TDPFs + integrate), evaluated ON THE HOST through the plugin's self-check hook and
    integrated with the reference's own solver call (scipy RK45, base.py:855), reproduces the unmodified reference."""
    import ctypes as C
    from datetime import datetime
    from scipy.integrate import solve_ivp
    g = golden["tdpf"]
    num, date, ramp = gm.make_tdpf_models()

    def integrate_on_host(model, par, n_out, rtol=1e-4):
        fn = model.device_model.lib.sodm_plugin_host_rhs
        fn.restype, fn.argtypes = None, [C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        buf = np.concatenate([par, [0.0]])

        def rhs(t, y):
            dy = np.empty(3)
            yy = np.ascontiguousarray(y)
            fn(float(t), yy.ctypes.data, buf.ctypes.data, dy.ctypes.data)
            return dy
        return solve_ivp(rhs, [0, n_out - 1], [1000.0, 1.0, 0.0], t_eval=np.arange(n_out), rtol=rtol)
    num._bind_generated_functor(None, [None])
    r = integrate_on_host(num, np.array([0.35, 5.0]), 121)
    np.testing.assert_allclose(r.y, g["numeric_y"], rtol=1e-9, atol=1e-9)
    assert r.nfev == int(g["numeric_nfev"])
    start = datetime(2020, 1, 1)
    date._bind_generated_functor(86400e6, [start])
    assert date._date_params == ["start_measures"] and list(date._param_offsets) == ["beta", "gamma", "start_measures", "__t0__"]
    t0_abs = codegen.date_to_units(start, 86400e6)
    for i, shift in enumerate(g["date_shift_days"]):
        par = np.array([0.35, 5.0, t0_abs + 35 + shift, t0_abs])               # 2020-02-05 is day 35
        r = integrate_on_host(date, par, 121)
        np.testing.assert_allclose(r.y, g["date_y"][i], rtol=1e-9, atol=1e-9)
        assert r.nfev == int(g["date_nfev"][i])
    ramp._bind_generated_functor(86400e6, [start])
    r = integrate_on_host(ramp, np.array([0.35, 5.0, 14.0, t0_abs]), 121)
    np.testing.assert_allclose(r.y, g["ramp_y"], rtol=1e-9, atol=1e-9)
    assert r.nfev == int(g["ramp_nfev"])


@pytest.mark.parametrize("make", [gm.make_influenza, gm.make_sirsi, gm.make_seird, gm.make_lotka, lambda: gm.make_pfr(5), gm.make_seasonal,
                                  lambda: gm.GenSir({"S": 990, "I": 10}, {"beta": 0.9, "gamma": 0.2}),
                                  lambda: gm.make_sirsi(gm.GenOdeSirSi),
                                  lambda: gm.GenOdeSir({"S": 999, "I": 1}, {"beta": 0.35, "gamma": 5})])
def test_generated_models_build_and_pass_their_self_check(make):
    model = make()                                                # traces, compiles for sm_100a, evaluates once on the host
    plug = model.device_model
    assert isinstance(plug, plugin.Plugin) and os.path.dirname(plug.path) == plugin.PLUGIN_DIR
    n_state = sum(int(np.prod(s)) if s else 1 for s in model.state_shapes.values())
    assert plug.info.n_state == n_state and plug.info.n_param == model._n_param_flat
    # every symbol of include/sodm_plugin.h is exported
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "sodm_plugin.h")).read(), flags=re.S)
    raw = C.CDLL(plug.path)
    for name in set(re.findall(r"\b(sodm_plugin_[a-z0-9_]+)\s*\(", text)):
        assert hasattr(raw, name), name
    assert make().device_model is plug                            # cached: same source, same library


def test_generated_influenza_lists_the_reference_law():
    model = gm.make_influenza()
    plug, lay = model.device_model, model.device_model.layout
    assert (lay["n_draw"], lay["n_channel"], lay["tdpf_offset"], lay["tdpf_count"]) == (24, 24, 8, 16)
    from pysodm_b200 import tutorial_data as td
    y = td.influenza_flat_y0()
    par = td.influenza_flat_par()
    par[8:] = td.CONTACT_TABLE.reshape(3, 16)[0]
    for quirk in (True, False):
        n, p = plugin.host_propose(plug, y, par, 0.5, quirk)
        states = {k: np.asarray(v, dtype=np.float64) for k, v in td.INFLUENZA_INITIAL_STATES.items()}
        for k in td.INFLUENZA_STATES:
            states.setdefault(k, np.zeros(4))
        pr = dict(td.INFLUENZA_PARAMETERS, N=td.CONTACT_TABLE.reshape(3, 4, 4)[0])
        rates = type(model).compute_rates(0.0, **{k: states[k] for k in td.INFLUENZA_STATES}, **pr)
        for name, e, m, drawn, first in lay["groups"]:
            probs = ojump.transition_probabilities([np.asarray(r, dtype=np.float64)[e] for r in rates[name]], 0.5, quirk)
            cond = probs[:m] / np.cumsum(probs[::-1])[::-1][:m]             # sequential conditional binomials of Multinomial(p / sum p)
            assert np.allclose(p[first:first + drawn], cond[:drawn], rtol=3e-6), (name, e, quirk)
            assert (n[first:first + drawn] == int(states[name][e])).all()


def test_three_exits_and_bookkeeping_of_a_new_model():
    model = gm.make_seird()
    plug, lay = model.device_model, model.device_model.layout
    assert lay["n_draw"] == 3 * 6 and lay["n_channel"] == 18
    dep = [d for d in range(lay["n_draw"]) if (lay["dep_mask"] >> d) & 1]
    assert len(dep) == 6                                          # exits 2 and 3 of I, for each of the three ages
    # SSA pieces on the host: propensities = state * rate in the reference's channel order; one event of each channel
    y = np.concatenate([np.ravel(np.asarray(model.initial_states[k], dtype=np.float64)) for k in model.states_names])
    par = np.concatenate([np.ravel(np.asarray(model.parameters[k], dtype=np.float64)) for k in model._spec_params])
    a = np.empty(lay["n_channel"])
    y_i = y.astype(np.int32)
    plug.lib.sodm_plugin_host_propensities(y_i.ctypes.data, np.concatenate([par, [0.0]]).ctypes.data, a.ctypes.data, 0.0)
    kw = {k: np.asarray(model.initial_states[k], dtype=np.float64) for k in model.states_names}
    rates = gm.SEIRD.compute_rates(0.0, **kw, **model.parameters)
    want = np.concatenate([np.ravel(kw[s] * np.asarray(r, dtype=np.float64)) for s, lst in rates.items() for r in lst])
    assert np.allclose(a, want, rtol=1e-14)
    for ch in range(lay["n_channel"]):
        yy = y_i.copy()
        plug.lib.sodm_plugin_host_fire(yy.ctypes.data, ch)
        T = {s: [np.zeros(3) for _ in lst] for s, lst in rates.items()}
        (s, i, e), = [k for k, c in lay["channel_slots"].items() if c == ch]
        T[s][i][e] = 1
        new = gm.SEIRD.apply_transitionings(0.0, 1.0, T, **kw, **model.parameters)
        assert np.array_equal(yy, np.concatenate([np.ravel(v) for v in new]).astype(np.int32)), ch


def test_random_numpy_expressions_translate_exactly():
    """Fuzz: random compositions of the NumPy idioms model functions use (broadcasting, slices, matmul, reductions,
    stacking, ufuncs, branches) -- the compiled functor, evaluated on the host, agrees with NumPy to rounding."""
    import ctypes as C
    rng = np.random.default_rng(123)
    n = 4

    def build(seed):
        r = np.random.default_rng(seed)
        M = r.uniform(0.5, 2.0, size=(n, n))
        w = r.uniform(0.1, 1.0, size=n)
        k1, k2, k3 = r.integers(0, 6, size=3)

        def integrate(t, A, B, s, p, q):
            tot = A + B + s
            x = [A / tot, np.sqrt(A * B + 1.0), np.exp(-0.1 * A) * B, (M @ A) / np.sum(B), A[::-1] * w, np.tanh(0.01 * (A - B))][k1]
            y = [np.concatenate([A[1:], A[:1]]) - A, np.cumsum(B)[-1] * np.ones(n), np.stack([A, B]).sum(axis=0) * q,
                 np.where(A > B, A - B, 0.5 * (B - A)), np.maximum(A, 2.0 * B) - np.minimum(A, B), (A ** 2 + p) / (1.0 + B)][k2]
            if t < 3.0:
                z = p * s
            elif s > 1.0 and q < 5.0:
                z = np.log1p(s) + np.dot(w, x)
            else:
                z = -s / [1.0, 2.0, q][k3 % 3]
            dA = -p * x + y * 0.01
            dB = p * x - q * B * 0.001 + np.abs(np.sin(0.3 * t))
            return dA, dB, z - 0.1 * s
        return integrate
    shapes = {"A": (n,), "B": (n,), "s": ()}
    pars = {"p": (), "q": ()}
    for seed in range(12):
        fn = build(seed)
        src, n_state, n_param = codegen.ode_source(fn, shapes, pars)
        plug = plugin.compile_source(src, {"n_state": n_state})
        host = plug.lib.sodm_plugin_host_rhs
        host.restype, host.argtypes = None, [C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        for _ in range(6):
            A, B = rng.uniform(0.5, 20, size=n), rng.uniform(0.5, 20, size=n)
            s, p, q, t = rng.uniform(0.2, 3), rng.uniform(0.1, 1), rng.uniform(0.5, 8), rng.uniform(0, 6)
            y = np.concatenate([A, B, [s]])
            par = np.array([p, q, 0.0])
            dy = np.empty(n_state)
            host(float(t), y.ctypes.data, par.ctypes.data, dy.ctypes.data)
            want = np.concatenate([np.ravel(v) for v in fn(t, A, B, s, p, q)])
            np.testing.assert_allclose(dy, want, rtol=1e-12, atol=1e-13, err_msg=f"seed {seed}")


def test_large_stencil_models_take_the_warp_layout():
    """>= 64 equations that fall into a few expression shapes: the translation groups them by shape, pads every group to
    whole warps and emits one template per position; evaluated on the host, the templates reproduce the plain functor and
    the Python function (the self-check inside build_ode), for the packed-bed reactor and for a reaction-diffusion ring."""
    pfr = gm.make_pfr(25)
    assert pfr.device_model.layout.get("warp") and pfr.device_model.info.n_state == 200
    g, outs = codegen.trace_ode(type(pfr).integrate, {k: tuple(v) for k, v in pfr.state_shapes.items()},
                                {k: tuple(np.shape(pfr.parameters[k])) for k in pfr._spec_params})
    classes = codegen.warp_layout(g, outs)
    assert sorted(len(c["members"]) for c in classes) == [8, 48, 48, 96]        # inlet, solid (two signs), fluid incl. outlet

    class Ring(gm.ODE):                                     # two species on a ring of 40 cells: 80 equations, 2 shapes
        states, parameters, dimensions = ["u", "v"], ["D", "a", "b"], ["cell"]

        @staticmethod
        def integrate(t, u, v, D, a, b):
            lap = np.roll(u, 1) - 2 * u + np.roll(u, -1)
            return D * lap + a - (b + 1) * u + u ** 2 * v, b * u - u ** 2 * v + 0.1 * D * (np.roll(v, 1) - 2 * v + np.roll(v, -1))
    rng = np.random.default_rng(0)
    ring = Ring({"u": 1 + 0.1 * rng.random(40), "v": 3 + 0.1 * rng.random(40)}, {"D": 0.2, "a": 1.0, "b": 3.0},
                coordinates={"cell": list(range(40))})
    assert ring.device_model.layout.get("warp")

    class Irregular(gm.ODE):                                # 64 equations, every one of a different shape: stays thread-per-draw
        states, parameters, dimensions = ["x"], ["k"], ["i"]

        @staticmethod
        def integrate(t, x, k):
            out = np.zeros(64)
            for i in range(64):
                acc = -k * x[i]
                for j in range(i % 29 + 1):
                    acc = acc + 1e-3 * x[(i + j) % 64] * (j + 1)
                out[i] = acc
            return out,
    irr = Irregular({"x": np.ones(64)}, {"k": 0.1}, coordinates={"i": list(range(64))})
    assert not irr.device_model.layout.get("warp")


def test_equality_on_model_values_is_symbolic():
    """`==` / `!=` used to be node identity, so `if I == 0:` silently traced one branch.  Both are comparisons now."""
    from pysodm_b200.base import JumpProcess

    class EqSir(JumpProcess):
        states = ['S', 'I', 'R']
        parameters = ['beta', 'gamma']

        @staticmethod
        def compute_rates(t, S, I, R, beta, gamma):
            if I == 0:
                r = 0.0 * beta
            else:
                r = beta * (I / (S + I + R))
            g_ = gamma if t != 5 else 2 * gamma
            return {'S': [r], 'I': [g_ * np.ones(1)]}

        @staticmethod
        def apply_transitionings(t, tau, transitionings, S, I, R, beta, gamma):
            T = transitionings
            return S - T['S'][0], I + T['S'][0] - T['I'][0], R + T['I'][0]
    m = EqSir({'S': 990, 'I': 10}, {'beta': 0.9, 'gamma': 0.2})
    par = np.array([0.9, 0.2])
    q = lambda x: x / (1 + x)                                            # noqa: E731  (stay-probability quirk, single exit)
    _, p = plugin.host_propose(m.device_model, [990, 10, 0], par, 0.5, t=3.0)
    assert np.allclose(p, [q(1 - np.exp(-0.5 * 0.9 * 0.01)), q(1 - np.exp(-0.1))], rtol=2e-5)
    _, p = plugin.host_propose(m.device_model, [990, 0, 10], par, 0.5, t=5.0)       # I == 0 and t == 5: the other branches
    assert p[0] == 0.0 and np.isclose(p[1], q(1 - np.exp(-0.2)), rtol=2e-5)
