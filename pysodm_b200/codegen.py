"""Translate a user's Python model functions into a device functor: the step from "four hand-written models" to "any
pySODM model as written" (SURVEY section 8f-2).

pySODM models are Python callables on small NumPy arrays -- `integrate(t, *states, *parameters)` for an ODE
(`/root/reference/src/pySODM/models/base.py:746-796` calls it at every Runge-Kutta stage), `compute_rates(t, ...)` and
`apply_transitionings(t, tau, transitionings, ...)` for a jump process (`base.py:352-417`).  They cannot run on the
GPU, and there is no CPU fallback here.  Instead the function is TRACED ONCE with symbolic inputs:

  * every state / parameter element becomes a `Sym` leaf, wrapped in NumPy arrays of dtype=object with the shapes the
    model declares, so NumPy itself does the broadcasting, slicing, `@` / `np.matmul`, `np.sum`, stacking, transposes
    ... and calls `Sym.__add__`, `Sym.exp()` etc. per element (that is how NumPy evaluates ufuncs on object arrays);
  * Python control flow runs as usual (loops over `range(nx)`, `if` on shapes): what is recorded is the straight-line
    arithmetic of one evaluation, as a hash-consed expression DAG.  Branching on a VALUE (`if S > 0`, `np.where` on a
    state) cannot be traced and raises `TraceError` with the offending operation;
  * `np.zeros / ones / empty / full (_like)` are replaced, for the duration of the trace, by versions that return
    object arrays, so the common "allocate, then fill element by element" style works;
  * numba-jitted functions are traced through their `py_func`.

The DAG is emitted as straight-line C++ (`double` for ODE right-hand sides, `float` for transition rates, `int` for the
bookkeeping of `apply_transitionings`), wrapped in a functor with the interface of the hand-written ones
(`csrc/ode_models.cuh`, `csrc/jump_models.cuh`), compiled by nvcc for sm_100a into a plugin library next to
`libsodm.so` (`plugin.py`) and driven by the same kernels.  Operation ORDER is the Python function's own, so the result
differs from NumPy's by rounding only (FMA contraction; FP32 rates as in the hand-written functors).
"""
import math
import types

import numpy as np


class TraceError(NotImplementedError):
    pass


# ---------------------------------------------------------------------------------------------------------------------
# expression nodes
# ---------------------------------------------------------------------------------------------------------------------
class Graph:
    """Hash-consed expression DAG of one trace."""

    def __init__(self):
        self.nodes = []          # (op, args, kind) ; args = node ids, or payload for leaves
        self.index = {}

    def node(self, op, args, kind):
        key = (op, args)
        i = self.index.get(key)
        if i is None:
            i = len(self.nodes)
            self.nodes.append((op, args, kind))
            self.index[key] = i
        return Sym(self, i)

    def const(self, value):
        if isinstance(value, (bool, np.bool_)):
            value = int(value)
        if isinstance(value, (int, np.integer)):
            return self.node("const", (int(value), "i"), "i")
        value = float(value)
        # an integral float constant keeps integer arithmetic integral (1.0 * S): the values are the same
        if math.isfinite(value) and value == math.floor(value) and abs(value) < 2 ** 31:
            return self.node("const", (int(value), "f"), "i")
        return self.node("const", (value, "f"), "f")

    def leaf(self, name, kind):
        return self.node("in", (name,), kind)


_UNARY = ("exp", "log", "sqrt", "expm1", "log1p", "sin", "cos", "tan", "tanh", "arctan", "floor", "ceil", "log10", "log2",
          "sinh", "cosh", "arcsin", "arccos", "exp2", "cbrt")


class Sym:
    """One scalar of a traced evaluation."""
    __slots__ = ("g", "i")

    def __init__(self, g, i):
        self.g, self.i = g, i

    @property
    def kind(self):
        return self.g.nodes[self.i][2]

    def _lift(self, other):
        if isinstance(other, Sym):
            if other.g is not self.g:
                raise TraceError("expression mixes values of two different traces")
            return other
        if isinstance(other, (int, float, np.integer, np.floating, bool, np.bool_)):
            return self.g.const(other)
        return None

    def _bin(self, op, a, b):
        ia, ib = self.g.nodes[a.i], self.g.nodes[b.i]
        if ia[0] == "const" and ib[0] == "const":                       # constant folding
            x, y = ia[1][0], ib[1][0]
            try:
                val = {"add": x + y, "sub": x - y, "mul": x * y, "div": x / y if op == "div" else None,
                       "pow": x ** y if op == "pow" else None}[op]
                return self.g.const(val)
            except (ZeroDivisionError, OverflowError, TypeError):
                pass
        # exact identities only (x + 0, x - 0, x * 1, x / 1, x ** 1): no change of value, not even of rounding
        if ib[0] == "const" and ((op in ("add", "sub") and ib[1][0] == 0) or (op in ("mul", "div", "pow") and ib[1][0] == 1)):
            return a
        if ia[0] == "const" and ((op == "add" and ia[1][0] == 0) or (op == "mul" and ia[1][0] == 1)):
            return b
        kind = "i" if (op in ("add", "sub", "mul") and a.kind == "i" and b.kind == "i") else "f"
        return self.g.node(op, (a.i, b.i), kind)

    def _binary(self, op, other, swap=False):
        o = self._lift(other)
        if o is None:
            return NotImplemented
        return self._bin(op, o, self) if swap else self._bin(op, self, o)

    def __add__(self, o): return self._binary("add", o)
    def __radd__(self, o): return self._binary("add", o, True)
    def __sub__(self, o): return self._binary("sub", o)
    def __rsub__(self, o): return self._binary("sub", o, True)
    def __mul__(self, o): return self._binary("mul", o)
    def __rmul__(self, o): return self._binary("mul", o, True)
    def __truediv__(self, o): return self._binary("div", o)
    def __rtruediv__(self, o): return self._binary("div", o, True)
    def __pow__(self, o): return self._binary("pow", o)
    def __rpow__(self, o): return self._binary("pow", o, True)

    def __neg__(self):
        n = self.g.nodes[self.i]
        if n[0] == "const":
            return self.g.const(-n[1][0])
        return self.g.node("neg", (self.i,), self.kind)

    def __pos__(self):
        return self

    def __abs__(self):
        return self.g.node("abs", (self.i,), self.kind)

    def _unary(self, op):
        n = self.g.nodes[self.i]
        if n[0] == "const":
            return self.g.const(getattr(np, op)(float(n[1][0])))
        return self.g.node(op, (self.i,), "f")

    def _refuse(self, what):
        raise TraceError(f"{what} on a model value: the device translation records straight-line arithmetic and cannot "
                         "follow Python control flow that depends on states, parameters or time")

    def __bool__(self): self._refuse("truth test (if / while / and / or / np.where)")
    def __float__(self): self._refuse("float() (e.g. storing a model value into a float64 array made outside the function)")
    def __int__(self): self._refuse("int()")
    def __index__(self): self._refuse("use as an index")
    def __lt__(self, o): self._refuse("comparison")
    __le__ = __gt__ = __ge__ = __lt__

    def __eq__(self, o):
        return isinstance(o, Sym) and o.g is self.g and o.i == self.i

    def __hash__(self):
        return hash((id(self.g), self.i))

    def __repr__(self):
        return f"<Sym {self.i}:{self.g.nodes[self.i][0]}>"


for _name in _UNARY:                       # np.exp(object array) calls element.exp()
    setattr(Sym, _name, (lambda op: lambda self: self._unary(op))(_name))
Sym.conjugate = lambda self: self          # np.conj / some reductions on object arrays


def _sym_array(g, prefix, shape, kind, start):
    """Object array of leaves named (prefix, flat index) in C order, as the reference flattens states (base.py:464-467)."""
    n = int(np.prod(shape)) if shape else 1
    flat = np.empty(n, dtype=object)
    for e in range(n):
        flat[e] = g.leaf((prefix, start + e), kind)
    return flat.reshape(shape) if shape else flat[0]


# ---------------------------------------------------------------------------------------------------------------------
# tracing
# ---------------------------------------------------------------------------------------------------------------------
class _NumpyProxy:
    """`np` as seen by the traced function: array constructors return object arrays (so `out[j] = <model value>` works)."""

    def __init__(self, graph):
        self._g = graph

    def __getattr__(self, name):
        return getattr(np, name)

    def _obj(self, shape, fill):
        a = np.empty(shape, dtype=object)
        a.fill(fill)
        return a

    def zeros(self, shape, dtype=None, **kw): return self._obj(shape, 0.0)
    def ones(self, shape, dtype=None, **kw): return self._obj(shape, 1.0)
    def empty(self, shape, dtype=None, **kw): return self._obj(shape, 0.0)
    def full(self, shape, fill_value, dtype=None, **kw): return self._obj(shape, fill_value)
    def zeros_like(self, a, dtype=None, **kw): return self._obj(np.shape(a), 0.0)
    def ones_like(self, a, dtype=None, **kw): return self._obj(np.shape(a), 1.0)
    def empty_like(self, a, dtype=None, **kw): return self._obj(np.shape(a), 0.0)
    def full_like(self, a, fill_value, dtype=None, **kw): return self._obj(np.shape(a), fill_value)

    def where(self, *args, **kw):
        cond = np.asarray(args[0])
        if cond.dtype == object:
            raise TraceError("np.where on a model value cannot be translated (value-dependent selection)")
        return np.where(*args, **kw)

    def _select(self, op, a, b):
        a, b = np.asarray(a, dtype=object), np.asarray(b, dtype=object)
        out = np.empty(np.broadcast(a, b).shape, dtype=object)
        for idx, (x, y) in zip(np.ndindex(out.shape), zip(np.broadcast_to(a, out.shape).flat, np.broadcast_to(b, out.shape).flat)):
            xs = x if isinstance(x, Sym) else self._g.const(x)
            ys = y if isinstance(y, Sym) else self._g.const(y)
            out[idx] = self._g.node(op, (xs.i, ys.i), "i" if xs.kind == ys.kind == "i" else "f")
        return out if out.shape else out[()]

    def maximum(self, a, b, **kw): return self._select("max", a, b)
    def minimum(self, a, b, **kw): return self._select("min", a, b)


def _python_function(fn):
    fn = getattr(fn, "py_func", fn)                      # numba dispatcher -> the Python it was compiled from
    fn = getattr(fn, "__func__", fn)                     # staticmethod object
    if not isinstance(fn, types.FunctionType):
        raise TraceError(f"{fn!r} is not a Python function")
    return fn


def _retarget(fn, graph):
    """A copy of `fn` whose global `np` / `numpy` (under whatever alias) is the tracing proxy."""
    fn = _python_function(fn)
    g = dict(fn.__globals__)
    proxy = _NumpyProxy(graph)
    for k, v in list(g.items()):
        if v is np:
            g[k] = proxy
        elif getattr(v, "py_func", None) is not None and isinstance(getattr(v, "py_func"), types.FunctionType):
            g[k] = _retarget(v, graph)                   # jitted helpers called from the model function
    new = types.FunctionType(fn.__code__, g, fn.__name__, fn.__defaults__, fn.__closure__)
    new.__kwdefaults__ = fn.__kwdefaults__
    return new


def _collect(values, shapes, what):
    """Flatten the returned arrays (declaration order, C ravel) into a list of Sym / constants, checking shapes."""
    out = []
    if len(values) != len(shapes):
        raise TraceError(f"{what}: expected {len(shapes)} arrays ({list(shapes)}), the function returned {len(values)}")
    for (name, shape), v in zip(shapes.items(), values):
        a = np.asarray(v, dtype=object)
        want = tuple(shape)
        if a.shape != want:
            try:
                a = np.broadcast_to(a, want)
            except ValueError:
                raise TraceError(f"{what}: '{name}' has shape {a.shape}, expected {want}") from None
        out.extend(a.reshape(-1).tolist())
    return out


def trace_ode(integrate, state_shapes, param_shapes):
    """-> (graph, [derivative expression per flat state element])."""
    g = Graph()
    fn = _retarget(integrate, g)
    t = g.leaf(("t", 0), "f")
    args, off = [], 0
    for name, shape in state_shapes.items():
        args.append(_sym_array(g, "y", tuple(shape), "f", off))
        off += int(np.prod(shape)) if shape else 1
    off = 0
    for name, shape in param_shapes.items():
        args.append(_sym_array(g, "par", tuple(shape), "f", off))
        off += int(np.prod(shape)) if shape else 1
    res = fn(t, *args)
    if not isinstance(res, (tuple, list)):
        res = (res,)
    outs = _collect(res, state_shapes, "integrate")
    return g, [o if isinstance(o, Sym) else g.const(o) for o in outs]


def trace_jump(compute_rates, apply_transitionings, state_shapes, param_shapes):
    """-> dict with the rate graph (per state: list of transitions, each a flat list of expressions) and the bookkeeping
    graph (new value per flat state element as integer expressions of states `y` and transitionings `T`)."""
    state_sizes = {k: (int(np.prod(s)) if s else 1) for k, s in state_shapes.items()}
    # ---- rates: floating point
    g = Graph()
    fn = _retarget(compute_rates, g)
    t = g.leaf(("t", 0), "f")

    def inputs(graph, kind_y):
        args, off = [], 0
        for name, shape in state_shapes.items():
            args.append(_sym_array(graph, "y", tuple(shape), kind_y, off))
            off += state_sizes[name]
        off = 0
        for name, shape in param_shapes.items():
            args.append(_sym_array(graph, "par", tuple(shape), "f", off))
            off += int(np.prod(shape)) if shape else 1
        return args
    rates = fn(t, *inputs(g, "f"))
    if not isinstance(rates, dict):
        raise TraceError("compute_rates must return a dict {state: [rate arrays]}")
    channels = {}                                    # state -> list over transitions of flat expression lists
    for name, lst in rates.items():
        if name not in state_shapes:
            raise TraceError(f"compute_rates returned rates for '{name}', which is not a state")
        per = []
        for r in lst:
            a = np.asarray(r, dtype=object)
            if a.shape != tuple(state_shapes[name]):
                raise TraceError(f"rate of '{name}' has shape {a.shape}, expected {tuple(state_shapes[name])} (base.py:392-401)")
            per.append([x if isinstance(x, Sym) else g.const(x) for x in a.reshape(-1).tolist()])
        channels[name] = per
    # ---- bookkeeping: integers
    g2 = Graph()
    fn2 = _retarget(apply_transitionings, g2)
    t2, tau2 = g2.leaf(("t", 0), "f"), g2.leaf(("tau", 0), "f")
    T, c = {}, 0
    slots = {}                                       # (state, transition, element) -> channel number (reference order)
    for name in state_shapes:                        # channel order of `_SSA`: states in rate-dict order
        pass
    for name, per in channels.items():
        T[name] = []
        for i, flat in enumerate(per):
            arr = np.empty(len(flat), dtype=object)
            for e in range(len(flat)):
                arr[e] = g2.leaf(("T", c), "i")
                slots[(name, i, e)] = c
                c += 1
            T[name].append(arr.reshape(tuple(state_shapes[name])) if state_shapes[name] else arr[0])
    new = fn2(t2, tau2, T, *inputs(g2, "i"))
    if not isinstance(new, (tuple, list)):
        new = (new,)
    outs = _collect(new, state_shapes, "apply_transitionings")
    outs = [o if isinstance(o, Sym) else g2.const(o) for o in outs]
    for j, o in enumerate(outs):
        if o.kind != "i":
            raise TraceError("apply_transitionings must keep compartment counts integral (sums / differences / integer "
                             f"multiples of states and transitionings); flat state element {j} is a floating-point expression")
    return {"rate_graph": g, "channels": channels, "apply_graph": g2, "new_states": outs, "slots": slots,
            "n_channel": c, "state_sizes": state_sizes}


# ---------------------------------------------------------------------------------------------------------------------
# emission
# ---------------------------------------------------------------------------------------------------------------------
_CFUNC = {"exp": "exp", "log": "log", "sqrt": "sqrt", "expm1": "expm1", "log1p": "log1p", "sin": "sin", "cos": "cos",
          "tan": "tan", "tanh": "tanh", "arctan": "atan", "floor": "floor", "ceil": "ceil", "log10": "log10", "log2": "log2",
          "sinh": "sinh", "cosh": "cosh", "arcsin": "asin", "arccos": "acos", "exp2": "exp2", "cbrt": "cbrt", "abs": "fabs",
          "max": "fmax", "min": "fmin", "pow": "pow"}


def emit(graph, outputs, ctype, leaf_expr, prefix="v"):
    """Straight-line C++ for the nodes `outputs` depend on.  -> (list of statements, list of output expressions).
    ctype: 'double' | 'float' | 'int'.  leaf_expr(name) -> C expression of an input."""
    suffix = "f" if ctype == "float" else ""
    order, seen = [], set()
    stack = [(o.i, False) for o in reversed(outputs)]
    while stack:                                          # iterative post-order (graphs of 10^4 nodes)
        i, done = stack.pop()
        if done:
            order.append(i)
            continue
        if i in seen:
            continue
        seen.add(i)
        op, args, _ = graph.nodes[i]
        stack.append((i, True))
        if op not in ("const", "in"):
            for a in reversed(args):
                if a not in seen:
                    stack.append((a, False))
    name = {}
    lines = []

    def lit(value):
        if ctype == "int":
            return str(int(value))
        if isinstance(value, int):
            return f"{value}.0{suffix}"
        if math.isnan(value):
            return "NAN"
        if math.isinf(value):
            return ("-" if value < 0 else "") + "INFINITY"
        return repr(float(value)) + suffix
    for i in order:
        op, args, kind = graph.nodes[i]
        if op == "const":
            if ctype == "int" and kind != "i":
                raise TraceError("floating-point constant in integer bookkeeping")
            name[i] = lit(args[0])
            continue
        if op == "in":
            name[i] = leaf_expr(args[0])
            continue
        a = [name[x] for x in args]
        if op in ("add", "sub", "mul", "div"):
            if ctype == "int" and op == "div":
                raise TraceError("division in integer bookkeeping")
            expr = f"{a[0]} {'+-*/'['add sub mul div'.split().index(op)]} {a[1]}"
        elif op == "neg":
            expr = f"-{a[0]}"
        elif op == "pow":
            e = graph.nodes[args[1]]
            if e[0] == "const" and e[1][0] in (2, 3, 4) and isinstance(e[1][0], int):
                expr = " * ".join([a[0]] * e[1][0])                   # NumPy squares by multiplication too
            elif e[0] == "const" and e[1][0] == 0.5:
                expr = f"sqrt{suffix}({a[0]})"
            else:
                expr = f"pow{suffix}({a[0]}, {a[1]})"
        elif op == "abs" and ctype == "int":
            expr = f"abs({a[0]})"
        elif op in ("max", "min") and ctype == "int":
            expr = f"{op}({a[0]}, {a[1]})"
        else:
            if ctype == "int":
                raise TraceError(f"'{op}' in integer bookkeeping")
            expr = f"{_CFUNC[op]}{suffix}({', '.join(a)})"
        name[i] = f"{prefix}{i}"
        lines.append(f"const {ctype} {prefix}{i} = {expr};")
    return lines, [name[o.i] for o in outputs]
