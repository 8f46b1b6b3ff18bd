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
        self.forced, self.path = [], []          # branch exploration (see `explore`)
        self.unit_us = None                      # microseconds per simulation time unit when model functions see dates

    def decide(self, cond):
        """Truth value of a symbolic condition on the CURRENT exploration path: replay the forced prefix, then take the
        `True` side of every new condition (the driver comes back for the other side)."""
        for c, v in self.path:
            if c == cond.i:
                return v                                         # the same condition again: same answer
        k = len(self.path)
        if k < len(self.forced):
            if self.forced[k][0] != cond.i:
                raise TraceError("the model function is not deterministic under re-execution (branch order changed)")
            val = self.forced[k][1]
        else:
            val = True
        self.path.append((cond.i, val))
        return val

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
        raise TraceError(f"{what} on a model value cannot be translated: the device code is straight-line arithmetic plus "
                         "selections on comparisons (`if a < b`), a model value cannot become a Python number")

    def __bool__(self):
        n = self.g.nodes[self.i]
        if n[0] == "const":
            return bool(n[1][0])
        if n[2] != "b":
            self._refuse("truth test of a number")
        return self.g.decide(self)

    def __float__(self): self._refuse("float() (e.g. storing a model value into a float64 array made outside the function)")
    def __int__(self): self._refuse("int()")
    def __index__(self): self._refuse("use as an index")

    def _compare(self, op, other):
        o = self._lift(other)
        if o is None:
            if _is_datelike(other) or _is_deltalike(other):
                raise TraceError("a number is compared with a date: the model's functions use dates, so simulate over "
                                 "dates (time=['yyyy-mm-dd', 'yyyy-mm-dd'])")
            return NotImplemented
        a, b = self.g.nodes[self.i], self.g.nodes[o.i]
        if a[0] == "const" and b[0] == "const":
            x, y = a[1][0], b[1][0]
            return self.g.const(int({"lt": x < y, "le": x <= y, "gt": x > y, "ge": x >= y, "eq": x == y, "ne": x != y}[op]))
        return self.g.node(op, (self.i, o.i), "b")

    def __lt__(self, o): return self._compare("lt", o)
    def __le__(self, o): return self._compare("le", o)
    def __gt__(self, o): return self._compare("gt", o)
    def __ge__(self, o): return self._compare("ge", o)

    def _logic(self, op, other):
        o = self._lift(other)
        if o is None:
            return NotImplemented
        return self.g.node(op, (self.i, o.i), "b")

    def __and__(self, o): return self._logic("and", o)
    def __rand__(self, o): return self._logic("and", o)
    def __or__(self, o): return self._logic("or", o)
    def __ror__(self, o): return self._logic("or", o)
    def __invert__(self): return self.g.node("not", (self.i,), "b")

    # `==` / `!=` on model values are symbolic comparisons like `<` (as plain identity tests, `if S == 0:` silently
    # traced one branch only).  The SAME node compares equal as a Python bool, which is what dict / set membership of a Sym
    # needs (hash-consing: one node per expression); anything that cannot be a model value is simply "not equal".
    def __eq__(self, o):
        if isinstance(o, Sym) and o.g is self.g and o.i == self.i:
            return True
        r = self._compare("eq", o) if isinstance(o, (Sym, int, float, np.integer, np.floating, bool, np.bool_)) else NotImplemented
        return False if r is NotImplemented else r

    def __ne__(self, o):
        if isinstance(o, Sym) and o.g is self.g and o.i == self.i:
            return False
        r = self._compare("ne", o) if isinstance(o, (Sym, int, float, np.integer, np.floating, bool, np.bool_)) else NotImplemented
        return True if r is NotImplemented else r

    def __hash__(self):
        return hash((id(self.g), self.i))

    def __repr__(self):
        return f"<Sym {self.i}:{self.g.nodes[self.i][0]}>"

    # a model value behaves like a NumPy scalar / 0-d array where user code asks for its geometry
    shape, ndim, size = (), 0, 1

    @property
    def T(self): return self
    def sum(self, *a, **k): return self
    def copy(self): return self
    def item(self): return self
    def squeeze(self, *a, **k): return self
    def reshape(self, *shape):
        shape = shape[0] if len(shape) == 1 and isinstance(shape[0], (tuple, list)) else shape
        out = np.empty(shape, dtype=object)
        out.fill(self)
        if out.size != 1:
            raise ValueError(f"cannot reshape a scalar into shape {tuple(shape)}")
        return out
    def __getitem__(self, key):
        # the reference holds a dimensionless state as a ONE-element array (utils.py:45-55), so `states['I'][0]` is valid
        # user code there: element 0 of a scalar model value is the value itself
        if key is Ellipsis or (isinstance(key, tuple) and (key == () or key == (0,))) or (isinstance(key, (int, np.integer)) and key in (0, -1)):
            return self
        raise IndexError("a scalar model value cannot be indexed")


for _name in _UNARY:                       # np.exp(object array) calls element.exp()
    setattr(Sym, _name, (lambda op: lambda self: self._unary(op))(_name))
Sym.conjugate = lambda self: self          # np.conj / some reductions on object arrays


# ---------------------------------------------------------------------------------------------------------------------
# dates: what the reference hands to model functions when `sim` is given dates (base.py:759-764: `date = time_to_date(t)`)
# ---------------------------------------------------------------------------------------------------------------------
_EPOCH = np.datetime64("1970-01-01T00:00:00", "us")


def date_to_units(value, unit_us):
    """A datetime-like as ABSOLUTE time in simulation time units (time units since 1970-01-01)."""
    d = np.datetime64(value, "us") if not hasattr(value, "to_datetime64") else value.to_datetime64().astype("datetime64[us]")
    return float((d - _EPOCH) / np.timedelta64(1, "us")) / unit_us


def _is_datelike(x):
    return hasattr(x, "to_datetime64") or isinstance(x, np.datetime64) or type(x).__name__ in ("datetime", "date")


def _is_deltalike(x):
    return hasattr(x, "to_timedelta64") or isinstance(x, np.timedelta64) or type(x).__name__ == "timedelta"


def _delta_to_units(value, unit_us):
    d = value.to_timedelta64() if hasattr(value, "to_timedelta64") else np.timedelta64(value, "us") if not isinstance(value, np.timedelta64) else value
    return float(d / np.timedelta64(1, "us")) / unit_us


class SymDate:
    """A symbolic datetime: `abs` = a Sym holding absolute time in simulation units.  Supports what time-dependent
    parameter functions do with dates: ordering against datetimes / date-valued parameters, +- timedeltas, differences."""

    def __init__(self, abs_sym, unit_us):
        self.abs, self.unit_us = abs_sym, unit_us

    def _other(self, o):
        if isinstance(o, SymDate):
            return o.abs
        if _is_datelike(o) or isinstance(o, str):
            return date_to_units(o, self.unit_us)
        raise TraceError(f"cannot compare the simulation date with {type(o).__name__}")

    def __lt__(self, o): return self.abs < self._other(o)
    def __le__(self, o): return self.abs <= self._other(o)
    def __gt__(self, o): return self.abs > self._other(o)
    def __ge__(self, o): return self.abs >= self._other(o)

    def __add__(self, o):
        if isinstance(o, SymDelta):
            return SymDate(self.abs + o.units, self.unit_us)
        if _is_deltalike(o):
            return SymDate(self.abs + _delta_to_units(o, self.unit_us), self.unit_us)
        return NotImplemented
    __radd__ = __add__

    def __sub__(self, o):
        if isinstance(o, SymDelta):
            return SymDate(self.abs - o.units, self.unit_us)
        if _is_deltalike(o):
            return SymDate(self.abs - _delta_to_units(o, self.unit_us), self.unit_us)
        return SymDelta(self.abs - self._other(o), self.unit_us)

    def __rsub__(self, o):
        return SymDelta(self._other(o) - self.abs, self.unit_us)

    def __getattr__(self, name):
        raise TraceError(f"`.{name}` of the simulation date cannot be translated (calendar logic such as weekdays belongs in "
                         "a tabulated time-dependent parameter of a jump process, or must be expressed through date comparisons)")


class SymDelta:
    """Difference of two symbolic dates; `units` = a Sym in simulation time units."""

    def __init__(self, units, unit_us):
        self.units, self.unit_us = units, unit_us

    @property
    def days(self):
        return (self.units * (self.unit_us / 86400e6)).floor() if isinstance(self.units, Sym) else np.floor(self.units * self.unit_us / 86400e6)

    def total_seconds(self):
        return self.units * (self.unit_us / 1e6)

    def __truediv__(self, o):
        if _is_deltalike(o):
            return self.units / _delta_to_units(o, self.unit_us)
        if isinstance(o, SymDelta):
            return self.units / o.units
        return SymDelta(self.units / o, self.unit_us)

    def __mul__(self, o): return SymDelta(self.units * o, self.unit_us)
    __rmul__ = __mul__

    def __add__(self, o):
        if isinstance(o, SymDelta):
            return SymDelta(self.units + o.units, self.unit_us)
        if isinstance(o, SymDate):
            return SymDate(o.abs + self.units, self.unit_us)
        if _is_deltalike(o):
            return SymDelta(self.units + _delta_to_units(o, self.unit_us), self.unit_us)
        if _is_datelike(o):
            return SymDate(date_to_units(o, self.unit_us) + self.units, self.unit_us)
        return NotImplemented
    __radd__ = __add__

    def __rsub__(self, o):                                   # date - symbolic difference
        if _is_datelike(o):
            return SymDate(date_to_units(o, self.unit_us) - self.units, self.unit_us)
        if _is_deltalike(o):
            return SymDelta(_delta_to_units(o, self.unit_us) - self.units, self.unit_us)
        return NotImplemented

    def _other(self, o):
        if isinstance(o, SymDelta):
            return o.units
        if _is_deltalike(o):
            return _delta_to_units(o, self.unit_us)
        raise TraceError(f"cannot compare a time difference with {type(o).__name__}")

    def __lt__(self, o): return self.units < self._other(o)
    def __le__(self, o): return self.units <= self._other(o)
    def __gt__(self, o): return self.units > self._other(o)
    def __ge__(self, o): return self.units >= self._other(o)


_US_PER = {"weeks": 7 * 86400e6, "days": 86400e6, "hours": 3600e6, "minutes": 60e6, "seconds": 1e6, "milliseconds": 1e3,
           "microseconds": 1.0, "W": 7 * 86400e6, "D": 86400e6, "d": 86400e6, "h": 3600e6, "min": 60e6, "s": 1e6, "ms": 1e3, "us": 1.0}


class _PandasProxy:
    """`pd` as seen by a traced function: conversions to datetime pass symbolic dates through, a Timedelta of a model
    value (`pd.Timedelta(days=ramp_length)`) becomes a symbolic difference."""

    def __init__(self, real, graph):
        self._real, self._g = real, graph

    def Timedelta(self, value=None, unit=None, **kw):
        parts = list(kw.items()) + ([(unit or "ns", value)] if value is not None else [])
        if not any(isinstance(v, Sym) for _, v in parts):
            return self._real.Timedelta(value, unit, **kw) if value is not None else self._real.Timedelta(**kw)
        if self._g.unit_us is None:
            raise TraceError("pd.Timedelta of a model value needs a simulation over dates")
        total = 0.0
        for name, v in parts:
            if name not in _US_PER:
                raise TraceError(f"pd.Timedelta unit '{name}' is not supported in translated model functions")
            total = total + v * (_US_PER[name] / self._g.unit_us)
        return SymDelta(total, self._g.unit_us)

    def __getattr__(self, name):
        return getattr(self._real, name)

    def to_datetime(self, x, *a, **k):
        return x if isinstance(x, SymDate) else self._real.to_datetime(x, *a, **k)

    def Timestamp(self, x, *a, **k):
        return x if isinstance(x, SymDate) else self._real.Timestamp(x, *a, **k)


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
        if cond.dtype != object or len(args) != 3:
            return np.where(*args, **kw)
        a, b = np.asarray(args[1], dtype=object), np.asarray(args[2], dtype=object)
        shape = np.broadcast(cond, a, b).shape
        out = np.empty(shape, dtype=object)
        cb, ab, bb = (np.broadcast_to(x, shape) for x in (cond, a, b))
        for idx in np.ndindex(shape):
            out[idx] = select(self._g, cb[idx], ab[idx], bb[idx])
        return out if shape else out[()]

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


def select(g, cond, a, b):
    """`a if cond else b` as a node (constant conditions fold)."""
    a = a if isinstance(a, Sym) else g.const(a)
    b = b if isinstance(b, Sym) else g.const(b)
    if not isinstance(cond, Sym):
        return a if cond else b
    c = g.nodes[cond.i]
    if c[0] == "const":
        return a if c[1][0] else b
    if a.i == b.i:
        return a
    return g.node("select", (cond.i, a.i, b.i), "i" if a.kind == b.kind == "i" else ("b" if a.kind == b.kind == "b" else "f"))


MAX_PATHS = 256


def explore(g, run):
    """Execute `run()` (-> flat list of outputs) once per control-flow path of the model function and merge the results
    into `select` nodes: an `if` on a comparison of model values is followed down BOTH sides (first the true side, then,
    on re-execution with the decisions so far replayed, the false side), depth first."""
    results, forced = [], []
    while True:
        g.forced, g.path = forced, []
        outs = run()
        results.append((list(g.path), [o if isinstance(o, Sym) else g.const(o) for o in outs]))
        path = list(g.path)
        while path and path[-1][1] is False:
            path.pop()
        if not path:
            break
        forced = path[:-1] + [(path[-1][0], False)]
        if len(results) >= MAX_PATHS:
            raise TraceError(f"more than {MAX_PATHS} control-flow paths through the model function")
    g.forced, g.path = [], []

    def merge(group, depth):
        if len(group) == 1:
            return group[0][1]
        cond = Sym(g, group[0][0][depth][0])
        yes = [r for r in group if r[0][depth][1]]
        no = [r for r in group if not r[0][depth][1]]
        a, b = merge(yes, depth + 1), merge(no, depth + 1)
        if len(a) != len(b):
            raise TraceError("the branches of the model function return results of different sizes")
        return [select(g, cond, x, y) for x, y in zip(a, b)]
    return merge(results, 0)


def _python_function(fn):
    fn = getattr(fn, "py_func", fn)                      # numba dispatcher -> the Python it was compiled from
    fn = getattr(fn, "__func__", fn)                     # staticmethod object
    if not isinstance(fn, types.FunctionType):
        raise TraceError(f"{fn!r} is not a Python function")
    return fn


def _retarget(fn, graph):
    """A copy of `fn` whose global `np` / `numpy` (under whatever alias) is the tracing proxy."""
    if isinstance(fn, types.MethodType):                 # bound method (a TDPF living on a helper object): keep its `self`
        inner, owner = _retarget(fn.__func__, graph), fn.__self__
        return lambda *a, **k: inner(owner, *a, **k)
    fn = _python_function(fn)
    g = dict(fn.__globals__)
    proxy = _NumpyProxy(graph)
    for k, v in list(g.items()):
        if v is np:
            g[k] = proxy
        elif isinstance(v, types.ModuleType) and v.__name__ == "pandas":
            g[k] = _PandasProxy(v, graph)
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
                a = a.reshape(want) if a.size == max(1, int(np.prod(want))) else np.broadcast_to(a, want)
            except ValueError:
                raise TraceError(f"{what}: '{name}' has shape {a.shape}, expected {want}") from None
        out.extend(a.reshape(-1).tolist())
    return out


def _inputs(graph, state_shapes, param_shapes, kind_y):
    """Symbolic keyword arguments {state / parameter name: array of leaves}; leaves are numbered in flat order.  The
    model functions are called BY NAME, as the reference does (`self.compute_rates(date, **states, **params)`,
    base.py:392), so the order of the parameters in their signatures is free."""
    kw, off = {}, 0
    for name, shape in state_shapes.items():
        kw[name] = _sym_array(graph, "y", tuple(shape), kind_y, off)
        off += int(np.prod(shape)) if shape else 1
    off = 0
    for name, shape in param_shapes.items():
        kw[name] = _sym_array(graph, "par", tuple(shape), "f", off)
        off += int(np.prod(shape)) if shape else 1
    return kw


def trace_ode(integrate, state_shapes, param_shapes, tdpf=(), model_params=None, date_unit_us=None, date_params=()):
    """-> (graph, [derivative expression per flat state element]).

    tdpf: [(parameter name, function(t, states, param, **extras), [extra argument names])] -- evaluated in front of
    `integrate` at every call, with the ORIGINAL parameter values, exactly as the reference's `func` does (base.py:769-779);
    `model_params`: the names `integrate` takes (parameters only the TDPFs use are dropped before the call, :781-785).
    date_unit_us: None for numeric time; else model functions receive a symbolic DATE (`SymDate`) built from `t` plus the
    LAST flat parameter (the absolute time of t = 0, filled in by `sim`), and `date_params` arrive as dates too."""
    g = Graph()
    g.unit_us = date_unit_us
    fn = _retarget(integrate, g)
    tdpf = [(name, _retarget(f, g), list(extra)) for name, f, extra in tdpf]
    t = g.leaf(("t", 0), "f")
    model_params = list(model_params) if model_params is not None else [k for k in param_shapes if k != "__t0__"]

    def run():
        kw = _inputs(g, state_shapes, param_shapes, "f")
        states = {k: kw[k] for k in state_shapes}
        pars = {k: kw[k] for k in param_shapes if k != "__t0__"}
        when = t
        if date_unit_us is not None:
            when = SymDate(t + kw["__t0__"], date_unit_us)
            for k in date_params:
                pars[k] = SymDate(pars[k], date_unit_us)
        params = dict(pars)
        for name, f, extra in tdpf:             # extra arguments from the dict being updated, own value from the original (:773-774)
            params[name] = f(when, states, pars[name], **{k: params[k] for k in extra})
        res = fn(when, **states, **{k: params[k] for k in model_params})
        if not isinstance(res, (tuple, list)):
            res = (res,)
        return _collect(res, state_shapes, "integrate")
    return g, explore(g, run)


def tdpf_traceable(fn, extras, state_shapes, param_shapes, target, date_unit_us=None, date_params=()):
    """Can the time-dependent parameter function `fn(t, states, param, **extras)` be translated?  It is called on symbolic
    time / states / parameters exactly as `trace_jump` will call it; calendar logic, table look-ups keyed by the date and
    anything else that needs a concrete value raise, and the function is then TABULATED on the host instead (which in turn
    requires that it does not read the states: base._tabulate_tdpfs hands it a guard).  -> (ok, reason)"""
    g = Graph()
    g.unit_us = date_unit_us
    t = g.leaf(("t", 0), "f")
    try:
        f = _retarget(fn, g)
    except TraceError as exc:
        return False, str(exc)

    def run():
        kw = _inputs(g, state_shapes, param_shapes, "f")
        when = t
        if date_unit_us is not None:
            when = SymDate(t + kw["__t0__"], date_unit_us)
            for k in date_params:
                kw[k] = SymDate(kw[k], date_unit_us)
        res = f(when, {k: kw[k] for k in state_shapes}, kw[target], **{k: kw[k] for k in extras})
        a = np.asarray(res, dtype=object)
        want = tuple(param_shapes[target])
        if a.shape != want and a.size != max(1, int(np.prod(want))):
            a = np.broadcast_to(a, want)
        for v in a.reshape(-1).tolist():
            if not isinstance(v, (Sym, int, float, np.integer, np.floating, bool, np.bool_)):
                raise TraceError(f"returns a {type(v).__name__}")
        return a.reshape(-1).tolist()
    try:
        explore(g, run)
    except Exception as exc:  # noqa: BLE001  (anything the symbolic call cannot digest means: not translatable)
        return False, f"{type(exc).__name__}: {exc}"
    return True, ""


def trace_jump(compute_rates, apply_transitionings, state_shapes, param_shapes, date_unit_us=None, date_params=(), tdpf=(),
               model_params=None):
    """-> dict with the rate graph (per state: list of transitions, each a flat list of expressions) and the bookkeeping
    graph (new value per flat state element as integer expressions of states `y` and transitionings `T`).

    tdpf: [(parameter name, function(t, states, param, **extras), [extra argument names])] -- time-dependent parameter
    functions that are TRANSLATED (they may read the states, the time and per-trajectory parameters): evaluated in front
    of `compute_rates` and `apply_transitionings` at every step / event exactly as the reference's `func` does
    (base.py:376-380: own value from the original parameters, extra arguments from the dict being updated);
    `model_params`: the names the two model functions take (parameters only the TDPFs use are dropped, :386)."""
    state_sizes = {k: (int(np.prod(s)) if s else 1) for k, s in state_shapes.items()}
    # ---- rates: floating point
    g = Graph()
    g.unit_us = date_unit_us
    fn = _retarget(compute_rates, g)
    t = g.leaf(("t", 0), "f")
    model_params = list(model_params) if model_params is not None else [k for k in param_shapes if k != "__t0__"]

    structure = []                                   # [(state, number of transitions)] in rate-dict order

    def model_inputs(graph, kind_y, tt, strict=True):
        """(time argument, keyword arguments) as the reference passes them: a date when sim runs over dates."""
        kw = _inputs(graph, state_shapes, param_shapes, kind_y)
        when = tt
        if date_unit_us is not None:
            when = SymDate(tt + kw["__t0__"], date_unit_us)
            for k in date_params:
                kw[k] = SymDate(kw[k], date_unit_us)
        kw.pop("__t0__", None)
        if tdpf:
            states = {k: kw[k] for k in state_shapes}
            pars = {k: kw[k] for k in param_shapes if k != "__t0__"}
            params = dict(pars)
            for name, f, extra in tdpf:
                try:
                    params[name] = _retarget(f, graph)(when, states, pars[name], **{k: params[k] for k in extra})
                except TraceError:
                    if strict:
                        raise
                    # bookkeeping graph: integer states; a function that cannot run on them leaves the parameter as it is
                    # (apply_transitionings may not use parameters in a way that matters: its outputs must stay integral)
            kw = dict(states)
            kw.update({k: params[k] for k in model_params})
        return when, kw

    def run_rates():
        when, kw = model_inputs(g, "f", t)
        rates = fn(when, **kw)
        if not isinstance(rates, dict):
            raise TraceError("compute_rates must return a dict {state: [rate arrays]}")
        shape_now, flat = [], []
        for name, lst in rates.items():
            if name not in state_shapes:
                raise TraceError(f"compute_rates returned rates for '{name}', which is not a state")
            shape_now.append((name, len(lst)))
            for r in lst:
                a = np.asarray(r, dtype=object)
                want = tuple(state_shapes[name])
                if a.shape != want:
                    # the reference holds dimensionless states as one-element arrays, so np.array([gamma]) is a valid rate
                    # for a scalar state (src/tests/test_JumpProcess.py:19); anything else must broadcast to the state
                    try:
                        a = a.reshape(want) if a.size == max(1, int(np.prod(want))) else np.broadcast_to(a, want)
                    except ValueError:
                        raise TraceError(f"rate of '{name}' has shape {a.shape}, expected {want} (base.py:392-401)") from None
                flat.extend(a.reshape(-1).tolist())
        if structure and structure != shape_now:
            raise TraceError("compute_rates returns different transitions on different branches")
        structure[:] = shape_now
        return flat
    flat_rates = explore(g, run_rates)
    channels, k = {}, 0                              # state -> list over transitions of flat expression lists
    for name, m in structure:
        per = []
        for _ in range(m):
            per.append(flat_rates[k:k + state_sizes[name]])
            k += state_sizes[name]
        channels[name] = per
    # ---- bookkeeping: integers
    g2 = Graph()
    g2.unit_us = date_unit_us
    fn2 = _retarget(apply_transitionings, g2)
    t2, tau2 = g2.leaf(("t", 0), "f"), g2.leaf(("tau", 0), "f")
    T, c = {}, 0
    slots = {}                                       # (state, transition, element) -> channel number (reference order)
    for name, per in channels.items():
        T[name] = []
        for i, flat in enumerate(per):
            arr = np.empty(len(flat), dtype=object)
            for e in range(len(flat)):
                arr[e] = g2.leaf(("T", c), "i")
                slots[(name, i, e)] = c
                c += 1
            T[name].append(arr.reshape(tuple(state_shapes[name])) if state_shapes[name] else arr[0])
    def run_apply():
        when2, kw2 = model_inputs(g2, "i", t2, strict=False)
        new = fn2(when2, tau2, T, **kw2)
        if not isinstance(new, (tuple, list)):
            new = (new,)
        return _collect(new, state_shapes, "apply_transitionings")
    outs = explore(g2, run_apply)
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
            if op == "div" and ctype == "double":
                # same quotient bit for bit, but a zero numerator does not take the FP64 division's slow path (csrc/fdiv.cuh)
                expr = f"::sodm::div_z({a[0]}, {a[1]})"
            else:
                expr = f"{a[0]} {'+-*/'['add sub mul div'.split().index(op)]} {a[1]}"
        elif op in ("lt", "le", "gt", "ge", "eq", "ne"):
            lines.append(f"const bool {prefix}{i} = {a[0]} {dict(lt='<', le='<=', gt='>', ge='>=', eq='==', ne='!=')[op]} {a[1]};")
            name[i] = f"{prefix}{i}"
            continue
        elif op in ("and", "or", "not"):
            expr = f"!{a[0]}" if op == "not" else f"{a[0]} {'&&' if op == 'and' else '||'} {a[1]}"
            lines.append(f"const bool {prefix}{i} = {expr};")
            name[i] = f"{prefix}{i}"
            continue
        elif op == "select":
            lines.append(f"const {'bool' if kind == 'b' else ctype} {prefix}{i} = {a[0]} ? {a[1]} : {a[2]};")
            name[i] = f"{prefix}{i}"
            continue
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


# ---------------------------------------------------------------------------------------------------------------------
# functor sources
# ---------------------------------------------------------------------------------------------------------------------
def _indent(lines, n=8):
    pad = " " * n
    return "\n".join(pad + ln for ln in lines)


def _reject_time(graph, outputs, what):
    used = set()
    stack = [o.i for o in outputs]
    while stack:
        i = stack.pop()
        if i in used:
            continue
        used.add(i)
        op, args, _ = graph.nodes[i]
        if op == "in" and args[0][0] in ("t", "tau"):
            raise TraceError(f"{what} uses its `{args[0][0]}` argument explicitly; on the device, time dependence of a "
                             "jump process comes from time-dependent parameter functions")
        if op not in ("in", "const"):
            stack.extend(args)


def _uses_input(graph, outputs, prefix):
    seen, stack = set(), [o.i for o in outputs]
    while stack:
        i = stack.pop()
        if i in seen:
            continue
        seen.add(i)
        op, args, _ = graph.nodes[i]
        if op == "in":
            if args[0][0] == prefix:
                return True
        elif op != "const":
            stack.extend(args)
    return False


def ode_source(integrate, state_shapes, param_shapes, origin="", **trace_kw):
    """C++ translation unit of a plugin holding the ODE `integrate`.  -> (source, n_state, n_param)."""
    g, outs = trace_ode(integrate, state_shapes, param_shapes, **trace_kw)
    n_state = len(outs)
    n_param = sum(int(np.prod(s)) if s else 1 for s in param_shapes.values())

    def leaf(name):
        return {"t": "t", "y": f"y[{name[1]}]", "par": f"par[{name[1]}]"}[name[0]]
    lines, exprs = emit(g, outs, "double", leaf)
    body = _indent(lines) + "\n" + _indent([f"dy[{j}] = {e};" for j, e in enumerate(exprs)])
    # large systems: warp-per-draw layout when the equations fall into a few shapes (then the thread functor above only
    # serves the host self-check and its kernel is not instantiated)
    warp = ode_warp_source(g, outs, n_state, n_param) if n_state >= WARP_MIN_STATES else None
    entry = "SODM_ODE_PLUGIN(::sodm::GenModel)" if warp is None else warp[0] + "\nSODM_ODE_WARP_PLUGIN(::sodm::GenModel, ::sodm::GenWarp)"
    src = f"""// GENERATED by pysodm_b200.codegen from the Python function `integrate`{origin} -- do not edit.
// {len(g.nodes)} expression nodes, {len(lines)} statements.
#include "plugin.cuh"

namespace sodm {{
struct GenModel {{
    static constexpr int N = {n_state}, NPAR = {max(n_param, 1)};
    __host__ __device__ static void rhs(double t, const double* y, const double* par, double* dy) {{
        (void)t; (void)par;
{body}
    }}
}};
}}  // namespace sodm

{entry}

// one evaluation on the HOST (the functor is __host__ __device__): build-time self-check against the Python function
extern "C" void sodm_plugin_host_rhs(double t, const double* y, const double* par, double* dy) {{
    ::sodm::GenModel::rhs(t, y, par, dy);
}}
"""
    return src, n_state, n_param




def jump_source(compute_rates, apply_transitionings, state_shapes, param_shapes, tdpf_params=(), origin="",
                date_unit_us=None, date_params=(), tdpf=(), model_params=None):
    """C++ translation unit of a plugin holding a jump process.  `param_shapes` must list the time-dependent parameters
    LAST (they are read from the per-step table instead of the parameter vector).
    -> (source, layout) with layout = dict(n_state, n_param, n_channel, n_draw, tdpf_offset, tdpf_count, slot_of)."""
    names = list(param_shapes)
    n_tdpf = len(tdpf_params)
    if n_tdpf and names[-n_tdpf:] != list(tdpf_params):
        raise ValueError("time-dependent parameters must come last in the flat parameter order")
    tr = trace_jump(compute_rates, apply_transitionings, state_shapes, param_shapes, date_unit_us, date_params, tdpf=tdpf,
                    model_params=model_params)
    sizes = {k: (int(np.prod(s)) if s else 1) for k, s in param_shapes.items()}
    n_param = sum(sizes.values())
    tdpf_count = sum(sizes[k] for k in tdpf_params)
    tdpf_offset = n_param - tdpf_count if tdpf_count else 0
    n_state = sum(tr["state_sizes"].values())
    g, g2 = tr["rate_graph"], tr["apply_graph"]
    state_off, off = {}, 0
    for name, n in tr["state_sizes"].items():
        state_off[name] = off
        off += n
    # ---- which transitionings does the bookkeeping read?
    used_T = set()
    stack = [o.i for o in tr["new_states"]]
    seen = set()
    while stack:
        i = stack.pop()
        if i in seen:
            continue
        seen.add(i)
        op, args, _ = g2.nodes[i]
        if op == "in":
            if args[0][0] == "T":
                used_T.add(args[0][1])
        elif op != "const":
            stack.extend(args)
    _reject_time(g2, tr["new_states"], "apply_transitionings")
    # ---- draw slots: per compartment element, its exits in order; trailing exits nobody reads are not drawn (their
    # probability still enters the normaliser W of the reference's quirk, base.py:339-342)
    # Slot ORDER: element-major (element 0 of every state, then element 1, ...), i.e. one age group's compartments after
    # the other.  Measured on influenza (B200): 137 ms per 10^6 trajectories against 148 ms state-major -- a lane's cheap
    # and expensive draws then alternate, which keeps the lanes of a warp closer together (profiles/r1_sampler_ab.txt).
    slot_of, groups, n_draw, dep_mask = {}, [], 0, 0
    order = sorted(((e, si, name) for si, name in enumerate(tr["channels"]) for e in range(tr["state_sizes"][name])))
    for e, _, name in order:
        m = len(tr["channels"][name])
        drawn = m
        while drawn > 0 and tr["slots"][(name, drawn - 1, e)] not in used_T:
            drawn -= 1
        first = n_draw
        for i in range(drawn):
            slot_of[tr["slots"][(name, i, e)]] = n_draw
            if i > 0:
                dep_mask |= 1 << n_draw
            n_draw += 1
        groups.append((name, e, m, drawn, first))
    if n_draw > 64:
        raise TraceError(f"{n_draw} binomial draws per step; the sampler's work list holds at most 64")
    if n_draw == 0:
        raise TraceError("apply_transitionings reads none of the transitionings")
    all_rates = [x for name, per in tr["channels"].items() for flat in per for x in flat]
    uses_time = _uses_input(g, all_rates, "t")       # e.g. seasonal forcing written into compute_rates

    def leaf_f(name):
        if name[0] == "y":
            return f"(float)y[{name[1]}]"
        if name[0] == "t":
            return "(float)t"
        if name[0] == "par":
            e = name[1]
            return f"tab[{e - tdpf_offset}]" if tdpf_count and e >= tdpf_offset else f"pre.par[{e}]"
        raise TraceError(f"unexpected input {name}")

    def leaf_d(name):
        if name[0] == "y":
            return f"(double)y[{name[1]}]"
        if name[0] == "t":
            return "t"
        if name[0] == "par":
            return f"par[{name[1]}]"
        raise TraceError(f"unexpected input {name}")
    # Rates that depend neither on the state, nor on the time, nor on a tabulated parameter are constants of the trajectory:
    # their leave probabilities 1 - exp(-tau rate) are formed once, in `prepare` (as the hand-written functors do), instead of
    # one expm1 per exit and step.
    def varies(sym):
        seen, stack = set(), [sym.i]
        while stack:
            i = stack.pop()
            if i in seen:
                continue
            seen.add(i)
            op, args, _ = g.nodes[i]
            if op == "in":
                if args[0][0] in ("y", "t") or (args[0][0] == "par" and tdpf_count and args[0][1] >= tdpf_offset):
                    return True
            elif op != "const":
                stack.extend(args)
        return False
    keys = [(name, i, e) for name, per in tr["channels"].items() for i, flat in enumerate(per) for e in range(len(flat))]
    hoisted = [k for k, sym in zip(keys, all_rates) if not varies(sym)]
    hoist_index = {k: n for n, k in enumerate(hoisted)}
    step_rates = [sym for k, sym in zip(keys, all_rates) if k not in hoist_index]
    f_lines, f_exprs = emit(g, step_rates, "float", leaf_f)
    f_exprs = iter(f_exprs)
    d_lines, d_exprs = emit(g, all_rates, "double", leaf_d, prefix="u")
    rate_expr_d = dict(zip(keys, d_exprs))
    rate_expr = {k: (None if k in hoist_index else next(f_exprs)) for k in keys}
    h_lines, h_exprs = emit(g, [sym for k, sym in zip(keys, all_rates) if k in hoist_index], "float",
                            lambda nm: f"(float)par[{nm[1]}]", prefix="c")
    pre_lines = h_lines + [f"pre.p[{n}] = leave_prob({x}, tau);" for n, x in enumerate(h_exprs)]
    put = []
    for name, e, m, drawn, first in groups:
        j = state_off[name] + e
        ps = [f"pre.p[{hoist_index[(name, i, e)]}]" if (name, i, e) in hoist_index else f"leave_prob({rate_expr[(name, i, e)]}, tau)"
              for i in range(m)]
        put.append("{")
        for i in range(m):
            put.append(f"    const float p{i} = {ps[i]};")
        put.append(f"    const float W0 = quirk ? 1.0f + p{m - 1} : 1.0f;")
        for i in range(drawn):
            if i > 0:
                put.append(f"    const float W{i} = W{i - 1} - p{i - 1};")
            put.append(f"    w.put({first + i}, y[{j}], p{i} / W{i});")
        put.append("}")

    def leaf_i(name):
        if name[0] == "y":
            return f"y[{name[1]}]"
        if name[0] == "T":
            return f"T{name[1]}"
        raise TraceError("apply_transitionings may only combine states and transitionings")
    a_lines, a_exprs = emit(g2, tr["new_states"], "int", leaf_i, prefix="n")
    get = [f"const int T{c} = w.get({s});" for c, s in sorted(slot_of.items())]
    fire_T = [f"const int T{c} = ch == {c} ? 1 : 0;" for c in sorted(used_T)]
    assign = [f"y[{j}] = {e};" for j, e in enumerate(a_exprs)]
    prop = []
    for (name, i, e), c in sorted(tr["slots"].items(), key=lambda kv: kv[1]):
        prop.append(f"a[{c}] = (double)y[{state_off[name] + e}] * ({rate_expr_d[(name, i, e)]});")
    npar_arr = max(n_param, 1)
    src = f"""// GENERATED by pysodm_b200.codegen from the Python functions `compute_rates` / `apply_transitionings`{origin} -- do not edit.
// {n_state} flat states, {n_param} flat parameters ({tdpf_count} time dependent), {tr['n_channel']} channels, {n_draw} binomial draws per step.
#include "plugin.cuh"

namespace sodm {{
struct GenModel {{
    static constexpr int NS = {n_state}, NPAR = {npar_arr}, NCH = {tr['n_channel']}, TDPF_OFFSET = {tdpf_offset}, TDPF_COUNT = {tdpf_count};
    static constexpr int NDRAW = {n_draw};
    static constexpr uint64_t DEP_MASK = {dep_mask}ull;
    static constexpr bool USES_TIME = {'true' if uses_time else 'false'};
    struct Pre {{ float par[{npar_arr}]; float p[{max(len(hoisted), 1)}]; }};
    __host__ __device__ static Pre prepare(const double* par, float tau, bool) {{
        Pre pre;
        (void)tau;
#pragma unroll
        for (int e = 0; e < {npar_arr}; ++e) pre.par[e] = (float)par[e];
{_indent(pre_lines)}
        return pre;
    }}
    // rates in FP32 as in the hand-written functors; law of the draws: jump_models.cuh header (reference base.py:283-348)
    template <class W>
    __host__ __device__ static void propose(const int* y, const Pre& pre, const float* tab, float tau, bool quirk, W& w, double t = 0.0) {{
        (void)tab; (void)pre; (void)t;
{_indent(f_lines)}
{_indent(put)}
    }}
    template <class W>
    __host__ __device__ static void apply(int* y, const W& w) {{
{_indent(get)}
{_indent(a_lines)}
{_indent(assign)}
    }}
    // SSA: propensities at the event time t (ssa_trajectory hands it over when USES_TIME); tabulated time-dependent parameters
    // have been copied into par[TDPF_OFFSET ..] by the caller
    __host__ __device__ static void propensities(const int* y, const double* par, double* a, double t = 0.0) {{
        (void)par; (void)t;
{_indent(d_lines)}
{_indent(prop)}
    }}
    __host__ __device__ static void fire(int* y, int ch) {{
{_indent(fire_T)}
{_indent(a_lines)}
{_indent(assign)}
    }}
}};
}}  // namespace sodm

SODM_JUMP_PLUGIN(::sodm::GenModel)

// one evaluation on the HOST (the functor is __host__ __device__): build-time self-check against the Python functions
namespace {{
struct HostList {{
    int n[{n_draw}];
    float p[{n_draw}];
    __host__ __device__ void put(int d, int nd, float pd) {{ n[d] = nd; p[d] = pd; }}
    __host__ __device__ int get(int d) const {{ return n[d]; }}
}};
}}
extern "C" void sodm_plugin_host_propose(const int* y, const double* par, const float* tab, double tau, int quirk, double t, int* n, float* p) {{
    HostList w;
    const auto pre = ::sodm::GenModel::prepare(par, (float)tau, quirk != 0);
    ::sodm::GenModel::propose(y, pre, tab, (float)tau, quirk != 0, w, t);
    for (int d = 0; d < {n_draw}; ++d) {{ n[d] = w.n[d]; p[d] = w.p[d]; }}
}}
extern "C" void sodm_plugin_host_apply(int* y, const int* draws) {{
    HostList w;
    for (int d = 0; d < {n_draw}; ++d) w.n[d] = draws[d];
    ::sodm::GenModel::apply(y, w);
}}
extern "C" void sodm_plugin_host_propensities(const int* y, const double* par, double* a, double t) {{ ::sodm::GenModel::propensities(y, par, a, t); }}
extern "C" void sodm_plugin_host_fire(int* y, int ch) {{ ::sodm::GenModel::fire(y, ch); }}
"""
    layout = {"n_state": n_state, "n_param": n_param, "n_channel": tr["n_channel"], "n_draw": n_draw,
              "tdpf_offset": tdpf_offset, "tdpf_count": tdpf_count, "slot_of": slot_of, "groups": groups,
              "channel_slots": tr["slots"], "dep_mask": dep_mask, "uses_time": uses_time}
    return src, layout


# ---------------------------------------------------------------------------------------------------------------------
# warp-per-draw layout for large ODE systems: lanes evaluate ISOMORPHIC derivative expressions with lane-dependent operands
# ---------------------------------------------------------------------------------------------------------------------
# A stencil model (the packed-bed reactor: 8 x nx equations) has a handful of expression SHAPES -- interior fluid node,
# outlet node, solid node, inlet -- each instantiated at many (species, node) positions with different operands.  For the
# thread-per-draw functor that is 200 equations in one thread's registers; here the equations are grouped by shape, every
# group is padded to whole warps, and position k of lane l evaluates ONE template per k (same instructions for the whole
# warp) whose operands -- state elements, parameters, constants -- come from per-member index tables.  The state of a draw
# lives in shared memory between stages; the integrators of ode_run.cuh drive it through the same `Sys` concept as the
# hand-written reactor functor.
WARP_MIN_STATES = 64
_MAX_TREE = 4000


def _shape_and_leaves(graph, root):
    """Canonical shape string of the expression TREE under `root` (shared nodes expanded) and its leaves in DFS order."""
    leaves, parts, count = [], [], [0]

    def walk(i):
        count[0] += 1
        if count[0] > _MAX_TREE:
            raise TraceError("derivative expression too large for the warp layout")
        op, args, kind = graph.nodes[i]
        if op == "in":
            if args[0][0] == "t":
                parts.append("t")
            else:
                parts.append("L")
                leaves.append((args[0][0], args[0][1]))
            return
        if op == "const":
            parts.append("L")
            leaves.append(("c", float(args[0])))
            return
        parts.append(op + "(")
        for a in args:
            walk(a)
            parts.append(",")
        parts.append(")")
    walk(root.i)
    return "".join(parts), leaves


def warp_layout(graph, outs):
    """Group the outputs by shape.  -> list of classes {members: [flat state index], slots: [(kind, values per member)],
    position_of_leaf: [slot id per leaf position]}, in an order that puts large classes first."""
    by_shape = {}
    for j, o in enumerate(outs):
        shape, leaves = _shape_and_leaves(graph, o)
        by_shape.setdefault(shape, []).append((j, leaves))
    classes = []
    for shape, items in by_shape.items():
        members = [j for j, _ in items]
        n_leaf = len(items[0][1])
        slots, slot_of_pos, seen = [], [], {}
        for pos in range(n_leaf):
            kinds = {lv[pos][0] for _, lv in items}
            if len(kinds) != 1:
                # the same position is a state in one member and a parameter / constant in another: not one template
                return None
            kind = kinds.pop()
            values = tuple(lv[pos][1] for _, lv in items)
            key = (kind, values)
            if key not in seen:
                seen[key] = len(slots)
                slots.append((kind, values))
            slot_of_pos.append(seen[key])
        classes.append({"members": members, "slots": slots, "slot_of_pos": slot_of_pos, "root": outs[members[0]]})
    classes.sort(key=lambda c: -len(c["members"]))
    return classes


def _template_graph(graph, cls):
    """The representative member's expression rebuilt over slot leaves (hash-consing merges what repeats)."""
    g = Graph()
    pos = [0]

    def build(i):
        op, args, kind = graph.nodes[i]
        if op == "in" and args[0][0] == "t":
            return g.leaf(("t", 0), "f")
        if op in ("in", "const"):
            s = cls["slot_of_pos"][pos[0]]
            pos[0] += 1
            return g.leaf(("slot", s), "f")
        kids = [build(a) for a in args]
        return g.node(op, tuple(k.i for k in kids), kind)
    return g, build(cls["root"].i)


def ode_warp_source(graph, outs, n_state, n_param):
    """C++ of the warp-per-draw functor `GenWarp` (appended to the plugin's translation unit), or None when the model does
    not decompose into a few shapes.  -> (source, n_local)"""
    try:
        classes = warp_layout(graph, outs)
    except TraceError:
        return None
    if classes is None or len(classes) > 24:
        return None
    # positions: class after class, each padded to whole warps
    perm, k0 = [], []
    for c in classes:
        k0.append(len(perm) // 32)
        perm.extend(c["members"])
        perm.extend([-1] * (-len(c["members"]) % 32))
    n_local = len(perm) // 32
    if n_local > 32:
        return None
    tables, funcs = [], []
    for ci, c in enumerate(classes):
        g, root = _template_graph(graph, c)
        m = len(c["members"])
        decl = []

        def slot_expr(s, ci=ci, c=c, decl=decl, m=m):
            kind, values = c["slots"][s]
            same = all(v == values[0] for v in values)
            if kind == "c":
                if same:
                    v = values[0]
                    return ("NAN" if math.isnan(v) else ("-" if v < 0 else "") + "INFINITY") if not math.isfinite(v) else repr(float(v))
                name = f"kC{ci}_{s}"
                tables.append(f"SODM_BOTH_TABLES(double, {name}, {m}, {{{', '.join(repr(float(v)) for v in values)}}})")
                decl.append(f"const double s{s} = SODM_TAB({name})[r];")
                return f"s{s}"
            arr = "ys" if kind == "y" else "pr"
            if same:
                return f"{arr}[{int(values[0])}]"
            step = values[1] - values[0] if m > 1 else 0
            if all(values[q] == values[0] + q * step for q in range(m)):
                decl.append(f"const double s{s} = {arr}[{int(values[0])} + {int(step)} * r];")
                return f"s{s}"
            # two-level affine (species x node): index = a + b (r mod P) + c (r div P) -- no table either
            for P in range(2, m // 2 + 1):
                if m % P:
                    continue
                a0 = values[0]
                b1 = values[1] - values[0]
                c1 = values[P] - values[0]
                if all(values[q] == a0 + b1 * (q % P) + c1 * (q // P) for q in range(m)):
                    if f"const int q{P} = r / {P};" not in decl:
                        decl.insert(0, f"const int m{P} = r - q{P} * {P};")
                        decl.insert(0, f"const int q{P} = r / {P};")
                    decl.append(f"const double s{s} = {arr}[{int(a0)} + {int(b1)} * m{P} + {int(c1)} * q{P}];")
                    return f"s{s}"
            name = f"kI{ci}_{s}"
            tables.append(f"SODM_BOTH_TABLES(short, {name}, {m}, {{{', '.join(str(int(v)) for v in values)}}})")
            decl.append(f"const double s{s} = {arr}[SODM_TAB({name})[r]];")
            return f"s{s}"
        exprs = {}

        def leaf(name):
            if name[0] == "t":
                return "t"
            if name[1] not in exprs:
                exprs[name[1]] = slot_expr(name[1])
            return exprs[name[1]]
        lines, res = emit(g, [root], "double", leaf, prefix=f"w{ci}_")
        funcs.append(f"""    // shape {ci}: {m} equations, {len(c['slots'])} operand slots
    __host__ __device__ static __forceinline__ double shape{ci}(double t, const double* ys, const double* pr, int r) {{
        (void)t; (void)ys; (void)pr; (void)r;
{_indent(decl)}
{_indent(lines)}
        return {res[0]};
    }}""")
    body = []
    for k in range(n_local):
        ci = max(i for i in range(len(classes)) if k0[i] <= k)
        m = len(classes[ci]["members"])
        body.append(f"{{ const int r = {32 * (k - k0[ci])} + lane; dy[{k}] = r < {m} ? shape{ci}(t, ys, pr, r) : 0.0; }}")
    nl = "\n"
    src = f"""
// ---- warp-per-draw layout: {len(classes)} expression shapes over {n_state} equations, {n_local} positions per lane ----------------
// index / constant tables exist twice: in device memory for the kernels, in host memory for the build-time self-check
#ifdef __CUDA_ARCH__
#define SODM_TAB(name) name##_d
#else
#define SODM_TAB(name) name##_h
#endif
#define SODM_BOTH_TABLES(T, name, n, ...) static __device__ const T name##_d[n] = __VA_ARGS__; static const T name##_h[n] = __VA_ARGS__;
namespace sodm {{
{nl.join(tables)}
SODM_BOTH_TABLES(short, kPerm, {len(perm)}, {{{', '.join(str(p) for p in perm)}}})
struct GenWarp {{
    static constexpr int N = {n_state}, NPAR = {max(n_param, 1)}, NL = {n_local};
    __host__ __device__ static __forceinline__ int flat_index(int k, int lane) {{ return SODM_TAB(kPerm)[32 * k + lane]; }}
{nl.join(funcs)}
    // derivatives of this lane's positions from the draw's full state `ys` and parameters `pr`
    __host__ __device__ static __forceinline__ void rhs_lane(double t, const double* ys, const double* pr, int lane, double* dy) {{
{_indent(body)}
    }}
}};
}}  // namespace sodm

extern "C" void sodm_plugin_host_rhs_warp(double t, const double* y, const double* par, double* dy) {{
    for (int lane = 0; lane < 32; ++lane) {{
        double d[::sodm::GenWarp::NL];
        ::sodm::GenWarp::rhs_lane(t, y, par, lane, d);
        for (int k = 0; k < ::sodm::GenWarp::NL; ++k) {{
            const int j = ::sodm::GenWarp::flat_index(k, lane);
            if (j >= 0) dy[j] = d[k];
        }}
    }}
}}
"""
    return src, n_local
