"""oracle/theano_shim.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A stand-in for the part of Theano 0.8.2 (python-dependencies.txt:17; absent here, no network) that the
reference's Python layers use to build their forward/backward functions
(pylayers/pylayers/pylayers.py:29-41 SoftmaxLayer, :126-142 BalancedSeedLossLayer, :158-168
ConstrainLossLayer).  It plays the role oracle/eigen_shim plays for the C++ sources: with it the
reference's OWN graph-building statements run unmodified and in place (oracle/ref_layers.py pulls the
class bodies out of the reference with ``ast``), so the numpy restatements in loss_oracle.py are pinned
against the reference's code instead of against themselves.

Symbolic expressions are closures evaluated with ``torch`` on the CPU; ``T.grad`` is
``torch.autograd.grad`` of the cost closure.  ``theano.function`` evaluates in float32 (what
``T.ftensor4`` means) or, with ``set_dtype(torch.float64)``, in float64 for tight comparisons.
What the stand-in cannot pin is Theano's own float32 reduction order (unspecified).
"""
import numpy as np
import torch

_DTYPE = torch.float32


def set_dtype(dt):
    global _DTYPE
    _DTYPE = dt


def _lift(x):
    return x if isinstance(x, Sym) else Sym(lambda env, v=x: v)


class Sym(object):
    """A node of the expression graph: ``ev(env)`` returns a torch tensor (or a python scalar)."""
    __array_priority__ = 1000

    def __init__(self, ev):
        self.ev = ev

    # numpy ufuncs applied to a symbolic tensor (pylayers.py:33 writes np.exp(preds - preds_max))
    def __array_ufunc__(self, ufunc, method, *inputs, **kw):
        if method != "__call__":
            return NotImplemented
        table = {np.exp: exp, np.log: log, np.add: lambda a, b: _lift(a) + b, np.subtract: lambda a, b: _lift(a) - b,
                 np.multiply: lambda a, b: _lift(a) * b, np.true_divide: lambda a, b: _lift(a) / b,
                 np.negative: lambda a: -_lift(a)}
        if ufunc not in table:
            return NotImplemented
        return table[ufunc](*inputs)

    def _bin(self, other, op, swap=False):
        o = _lift(other)
        a, b = (o, self) if swap else (self, o)
        return Sym(lambda env: op(a.ev(env), b.ev(env)))

    def __add__(self, o): return self._bin(o, lambda x, y: x + y)
    def __radd__(self, o): return self._bin(o, lambda x, y: x + y, True)
    def __sub__(self, o): return self._bin(o, lambda x, y: x - y)
    def __rsub__(self, o): return self._bin(o, lambda x, y: x - y, True)
    def __mul__(self, o): return self._bin(o, lambda x, y: x * y)
    def __rmul__(self, o): return self._bin(o, lambda x, y: x * y, True)
    def __truediv__(self, o): return self._bin(o, lambda x, y: x / y)
    def __rtruediv__(self, o): return self._bin(o, lambda x, y: x / y, True)
    __div__ = __truediv__
    __rdiv__ = __rtruediv__

    def __neg__(self):
        return Sym(lambda env: -self.ev(env))

    def __getitem__(self, idx):
        return Sym(lambda env: self.ev(env)[idx])


class Placeholder(Sym):
    def __init__(self):
        Sym.__init__(self, lambda env: env[id(self)])


def ftensor4():
    return Placeholder()


def _axes(axis):
    return None if axis is None else (tuple(axis) if isinstance(axis, (tuple, list)) else (axis,))


def sum(x, axis=None, keepdims=False):  # noqa: A001 (Theano's name)
    x = _lift(x)
    ax = _axes(axis)
    return Sym(lambda env: torch.sum(x.ev(env)) if ax is None else torch.sum(x.ev(env), dim=ax, keepdim=keepdims))


def mean(x, axis=None, keepdims=False):
    x = _lift(x)
    ax = _axes(axis)
    return Sym(lambda env: torch.mean(x.ev(env)) if ax is None else torch.mean(x.ev(env), dim=ax, keepdim=keepdims))


def max(x, axis=None, keepdims=False):  # noqa: A001
    x = _lift(x)
    ax = _axes(axis)
    return Sym(lambda env: torch.max(x.ev(env)) if ax is None else torch.amax(x.ev(env), dim=ax, keepdim=keepdims))


def addbroadcast(x, *axes):
    return x


def log(x):
    x = _lift(x)
    return Sym(lambda env: torch.log(x.ev(env)))


def exp(x):
    x = _lift(x)
    return Sym(lambda env: torch.exp(x.ev(env)))


def maximum(a, b):
    a, b = _lift(a), _lift(b)

    def ev(env):
        x, y = a.ev(env), b.ev(env)
        if not torch.is_tensor(y):
            y = torch.as_tensor(y, dtype=x.dtype)
        if not torch.is_tensor(x):
            x = torch.as_tensor(x, dtype=y.dtype)
        return torch.maximum(x, y)
    return Sym(ev)


def clip(x, lo, hi):
    x = _lift(x)
    return Sym(lambda env: torch.clamp(x.ev(env), lo, hi))


def grad(cost, wrt):
    """T.grad(cost, wrt): symbolic gradient(s); wrt is a placeholder or a list of placeholders."""
    many = isinstance(wrt, (list, tuple))
    ws = list(wrt) if many else [wrt]

    def ev_all(env):
        env2 = dict(env)
        leaves = []
        for w in ws:
            t = env[id(w)].detach().clone().requires_grad_(True)
            env2[id(w)] = t
            leaves.append(t)
        c = cost.ev(env2)
        return torch.autograd.grad(c, leaves, allow_unused=True)

    outs = [Sym(lambda env, k=k: ev_all(env)[k]) for k in range(len(ws))]
    return outs if many else outs[0]


class _Function(object):
    def __init__(self, inputs, outputs):
        self.inputs, self.outputs = inputs, outputs

    def __call__(self, *args):
        assert len(args) == len(self.inputs)
        env = {id(p): torch.as_tensor(np.asarray(a), dtype=_DTYPE) for p, a in zip(self.inputs, args)}
        many = isinstance(self.outputs, (list, tuple))
        outs = [o.ev(env) for o in (self.outputs if many else [self.outputs])]
        outs = [np.asarray(o.detach().numpy() if torch.is_tensor(o) else o) for o in outs]
        return outs if many else outs[0]


def function(inputs, outputs):
    return _Function(inputs, outputs)


def install():
    """Return (theano, T) module stand-ins with the attributes the reference's layers touch."""
    import types
    T = types.ModuleType("theano.tensor")
    for name in ("ftensor4", "sum", "mean", "max", "addbroadcast", "log", "exp", "maximum", "clip", "grad"):
        setattr(T, name, globals()[name])
    theano = types.ModuleType("theano")
    theano.function = function
    theano.tensor = T
    return theano, T
