"""The slice of `Theano==1.0.5` MIND's planner touches: `planners/mind/trajectory_tree.py:153-177` builds
the 6-state kinematic bicycle model from `T.dscalar` symbols with `+ * /`, `T.cos / T.sin / T.tan` and `T.stack`;
`planners/ilqr/autodiff.py` differentiates it with `T.grad` and turns the expressions into callables with
`theano.function`; `planners/ilqr/dynamics.py:146-208` slices the stacked Jacobian (`J[:, :x_dim]`).

This module is a small symbolic engine for exactly that: scalar expression nodes with analytic derivatives, object
arrays of them for stacked tensors, and a code generator that compiles an expression (array) into one Python function
over `math` (no graph optimiser, no C compilation).  The Jacobians it produces are exact derivatives, checked against
finite differences in tests/test_front_end_cpu.py.  Anything else of Theano raises NotImplementedError.
"""
import math
import types

import numpy as np


class Expr:
    __array_ufunc__ = None            # numpy scalars defer to our reflected operators

    def __add__(self, o): return _add(self, _wrap(o))
    def __radd__(self, o): return _add(_wrap(o), self)
    def __sub__(self, o): return _add(self, _neg(_wrap(o)))
    def __rsub__(self, o): return _add(_wrap(o), _neg(self))
    def __mul__(self, o): return _mul(self, _wrap(o))
    def __rmul__(self, o): return _mul(_wrap(o), self)
    def __truediv__(self, o): return _div(self, _wrap(o))
    def __rtruediv__(self, o): return _div(_wrap(o), self)
    def __neg__(self): return _neg(self)
    def __pow__(self, p):
        if not isinstance(p, (int, float)) or int(p) != p or p < 0:
            raise NotImplementedError("theano_lite: only non-negative integer powers")
        out = Const(1.0)
        for _ in range(int(p)):
            out = _mul(out, self)
        return out


class Const(Expr):
    def __init__(self, v): self.v = float(v)


class Var(Expr):
    def __init__(self, name): self.name = name


class Op(Expr):
    def __init__(self, kind, *args): self.kind, self.args = kind, args


def _wrap(o):
    if isinstance(o, Expr):
        return o
    if isinstance(o, (int, float, np.integer, np.floating)):
        return Const(o)
    raise NotImplementedError("theano_lite: cannot mix %r into a scalar expression" % type(o))


def _is(c, v): return isinstance(c, Const) and c.v == v


def _add(a, b):
    if _is(a, 0.0): return b
    if _is(b, 0.0): return a
    if isinstance(a, Const) and isinstance(b, Const): return Const(a.v + b.v)
    return Op("add", a, b)


def _neg(a):
    if isinstance(a, Const): return Const(-a.v)
    return Op("neg", a)


def _mul(a, b):
    if _is(a, 0.0) or _is(b, 0.0): return Const(0.0)
    if _is(a, 1.0): return b
    if _is(b, 1.0): return a
    if isinstance(a, Const) and isinstance(b, Const): return Const(a.v * b.v)
    return Op("mul", a, b)


def _div(a, b):
    if _is(a, 0.0): return Const(0.0)
    if _is(b, 1.0): return a
    if isinstance(a, Const) and isinstance(b, Const): return Const(a.v / b.v)
    return Op("div", a, b)


def cos(x): return Op("cos", _wrap(x))
def sin(x): return Op("sin", _wrap(x))
def tan(x): return Op("tan", _wrap(x))


def _d(e, v):
    """d e / d v for one scalar variable v"""
    if isinstance(e, Const): return Const(0.0)
    if isinstance(e, Var): return Const(1.0 if e is v else 0.0)
    k, a = e.kind, e.args
    if k == "add": return _add(_d(a[0], v), _d(a[1], v))
    if k == "neg": return _neg(_d(a[0], v))
    if k == "mul": return _add(_mul(_d(a[0], v), a[1]), _mul(a[0], _d(a[1], v)))
    if k == "div": return _div(_add(_mul(_d(a[0], v), a[1]), _neg(_mul(a[0], _d(a[1], v)))), _mul(a[1], a[1]))
    if k == "cos": return _mul(_neg(sin(a[0])), _d(a[0], v))
    if k == "sin": return _mul(cos(a[0]), _d(a[0], v))
    if k == "tan": return _div(_d(a[0], v), _mul(cos(a[0]), cos(a[0])))
    raise NotImplementedError(k)


class SymArray:
    """stacked expressions: an object ndarray with numpy indexing (what `J[:, :x_dim]` / `expr[i]` need)"""

    def __init__(self, a): self.a = a
    @property
    def shape(self): return self.a.shape
    def __len__(self): return len(self.a)
    def __getitem__(self, idx):
        r = self.a[idx]
        return SymArray(r) if isinstance(r, np.ndarray) else r


def stack(items, axis=0):
    if axis != 0:
        raise NotImplementedError("theano_lite: stack along axis 0 only")
    rows = []
    for it in items:
        if isinstance(it, SymArray):
            rows.append(it.a)
        elif isinstance(it, (list, tuple)):
            rows.append(stack(list(it)).a)
        else:
            rows.append(_wrap(it))
    out = np.empty((len(rows),) + (rows[0].shape if isinstance(rows[0], np.ndarray) else ()), dtype=object)
    for i, r in enumerate(rows):
        out[i] = r
    return SymArray(out)


def grad(cost=None, wrt=None, disconnected_inputs="raise", known_grads=None, **kw):
    if cost is None or known_grads is not None:
        raise NotImplementedError("theano_lite: grad(cost=None, known_grads=...) (batched Jacobians) is not supported")
    if isinstance(cost, SymArray):
        raise NotImplementedError("theano_lite: grad of a non-scalar expression")
    if isinstance(wrt, (list, tuple)):
        return [_d(_wrap(cost), v) for v in wrt]
    return _d(_wrap(cost), wrt)


def dscalar(name=None): return Var(name or "v")


def _emit(e, names, memo, lines):
    """common-subexpression-aware code generation: returns the Python expression string of node e"""
    if isinstance(e, Const): return repr(e.v)
    if isinstance(e, Var): return names[id(e)]
    if id(e) in memo: return memo[id(e)]
    a = [_emit(x, names, memo, lines) for x in e.args]
    k = e.kind
    s = {"add": "(%s + %s)", "mul": "(%s * %s)", "div": "(%s / %s)", "neg": "(-%s)", "cos": "_cos(%s)", "sin": "_sin(%s)",
         "tan": "_tan(%s)"}[k] % tuple(a)
    if k in ("cos", "sin", "tan"):
        t = "t%d" % len(lines)
        lines.append("    %s = %s" % (t, s))
        s = t
    memo[id(e)] = s
    return s


def function(inputs, outputs, on_unused_input="raise", name=None, **kw):
    """theano.function: compile `outputs` (a scalar expression or a SymArray) into f(*inputs) -> float / ndarray"""
    names = {}
    for i, v in enumerate(inputs):
        if not isinstance(v, Var):
            raise NotImplementedError("theano_lite: function inputs must be dscalar symbols")
        names[id(v)] = "a%d" % i
    lines, memo = [], {}
    if isinstance(outputs, SymArray):
        flat = [_emit(_wrap(e), names, memo, lines) for e in outputs.a.reshape(-1)]
        ret = "_np.array([%s], dtype=_np.float64).reshape(%r)" % (", ".join(flat), tuple(outputs.shape))
    else:
        ret = _emit(_wrap(outputs), names, memo, lines)
    src = "def _f(%s):\n%s\n    return %s\n" % (", ".join("a%d" % i for i in range(len(inputs))), "\n".join(lines) or "    pass", ret)
    ns = {"_cos": math.cos, "_sin": math.sin, "_tan": math.tan, "_np": np}
    exec(compile(src, "<theano_lite:%s>" % (name or "f"), "exec"), ns)
    f = ns["_f"]
    f.source = src
    return f


class TensorVariable(Expr):       # only used by the reference in isinstance() tests of code paths MIND never takes
    pass


def _unsupported(name):
    def f(*a, **k):
        raise NotImplementedError("theano_lite: %s is not part of the supported slice" % name)
    return f


tensor = types.ModuleType("theano.tensor")
for _n, _v in dict(dscalar=dscalar, stack=stack, cos=cos, sin=sin, tan=tan, grad=grad, TensorVariable=TensorVariable).items():
    setattr(tensor, _n, _v)
for _n in ("dvector", "dmatrix", "tile", "identity_like", "dot", "exp", "log", "sqrt"):
    setattr(tensor, _n, _unsupported("theano.tensor." + _n))
