"""A NumPy-backed stand-in for the slice of the ``jax`` API the reference uses, so that the
reference's own source (``/root/reference/go_melt/computeFunctions.py``) can be *executed
unmodified* in a container without JAX and produce golden vectors for the oracle.

Test infrastructure only.  It exists to pin ``oracle/`` against the reference's real code:
``make_golden.py`` imports the reference through this shim and writes ``*.npz`` fixtures; the
fixtures (not the shim, not the reference) travel with the repo.

Emulated semantics
  * x64 disabled: every float64 result is rounded to float32, every int64 to int32, like
    ``jax_enable_x64=False`` (the reference never enables it, cF:15);
  * functional updates ``x.at[idx].set/add``;
  * ``jax.jit`` (identity, accepts static_argnames/static_argnums), ``jax.vmap`` (a Python loop
    over the leading axis + stack; in_axes=0 only - all the reference uses), ``jax.lax.scan``
    (Python loop + stacked ys), ``jax.lax.select``, ``jax.experimental.sparse.BCOO`` with ``@``;
  * ``pyevtk.hl.gridToVTK`` is a no-op stub (the reference imports it at module level, cF:12).

Not emulated: XLA's summation order inside reductions / scatter-adds (NumPy pairwise or
float64 accumulation instead), FMA contraction.  Those differ from real XLA at the 1e-7
relative level, which is the pin tolerance the golden tests use for float fields.
"""
import functools
import sys
import types

import numpy as np


def _down(x):
    """x64-disabled dtype policy."""
    if isinstance(x, np.ndarray) or isinstance(x, np.generic):
        if x.dtype == np.float64:
            x = x.astype(np.float32)
        elif x.dtype == np.int64:
            x = x.astype(np.int32)
        elif x.dtype == np.uint64:
            x = x.astype(np.uint32)
    return x


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


def _raw(x):
    return x.view(np.ndarray) if isinstance(x, JArr) else x


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def _idx(self):
        i = self.idx
        if isinstance(i, tuple):
            return tuple(_raw(np.asarray(v)) if isinstance(v, (np.ndarray, list)) else v for v in i)
        return _raw(np.asarray(i)) if isinstance(i, (np.ndarray, list)) else i

    def set(self, val):
        out = np.array(_raw(self.arr), copy=True)
        out[self._idx()] = np.asarray(_raw(val)).astype(out.dtype) if np.ndim(val) else val
        return _wrap(out)

    def add(self, val):
        out = np.array(_raw(self.arr), copy=True)
        np.add.at(out, self._idx(), np.asarray(_raw(val)).astype(out.dtype) if np.ndim(val) else val)
        return _wrap(out)

    def get(self):
        return _wrap(_raw(self.arr)[self._idx()])


class JArr(np.ndarray):
    """ndarray whose every ufunc / function result is cast by the x64-disabled policy."""

    __array_priority__ = 100

    @property
    def at(self):
        return _At(self)

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kw):
        ins = tuple(_raw(i) for i in inputs)
        if out is not None:
            kw["out"] = tuple(_raw(o) for o in out)
        res = getattr(ufunc, method)(*ins, **kw)
        return _wrap_tree(res)

    def __array_function__(self, func, types_, args, kwargs):
        args = _tree_map(_raw, args)
        kwargs = _tree_map(_raw, kwargs)
        return _wrap_tree(func(*args, **kwargs))

    def __getitem__(self, idx):
        idx = _tree_map(_raw, idx)
        return _wrap(_raw(self)[idx])

    def __iter__(self):
        base = _raw(self)
        for i in range(base.shape[0]):
            yield _wrap(base[i])

    def block_until_ready(self):
        return self

    def reshape(self, *a, **k):
        return _wrap(_raw(self).reshape(*a, **k))

    def astype(self, dt, **k):
        return _wrap(_raw(self).astype(dt, **k), cast=False) if np.dtype(dt) not in (np.float64, np.int64) \
            else _wrap(_raw(self).astype(dt, **k))

    def sum(self, *a, **k):
        return _wrap(_raw(self).sum(*a, **k))

    def mean(self, *a, **k):
        return _wrap(_raw(self).mean(*a, **k))

    def max(self, *a, **k):
        return _wrap(_raw(self).max(*a, **k))

    def min(self, *a, **k):
        return _wrap(_raw(self).min(*a, **k))

    def any(self, *a, **k):
        return _raw(self).any(*a, **k)

    def all(self, *a, **k):
        return _raw(self).all(*a, **k)

    def __deepcopy__(self, memo):
        return _wrap(np.array(_raw(self), copy=True))

    def __reduce__(self):  # pickle / dill as a plain ndarray
        return (np.asarray, (np.array(_raw(self)),))


def _tree_map(f, x):
    if isinstance(x, tuple):
        return tuple(_tree_map(f, v) for v in x)
    if isinstance(x, list):
        return [_tree_map(f, v) for v in x]
    if isinstance(x, dict):
        return {k: _tree_map(f, v) for k, v in x.items()}
    return f(x)


def _wrap(x, cast=True):
    if isinstance(x, (np.ndarray, np.generic)):
        if cast:
            x = _down(x)
        return np.asarray(x).view(JArr)
    return x


def _wrap_tree(x):
    return _tree_map(_wrap, x)


def _lift(fn):
    @functools.wraps(fn)
    def g(*a, **k):
        a = _tree_map(_raw, a)
        k = _tree_map(_raw, k)
        return _wrap_tree(fn(*a, **k))

    return g


def _array(x, dtype=None, **k):
    x = _tree_map(_raw, x)
    a = np.array(x, dtype=dtype) if dtype is not None else np.array(x)
    return _wrap(a, cast=dtype is None or np.dtype(dtype) in (np.float64, np.int64))


def _creator(fn):
    def g(*a, dtype=None, **k):
        a = _tree_map(_raw, a)
        r = fn(*a, **k) if dtype is None else fn(*a, dtype=dtype, **k)
        return _wrap(r)

    return g


def _bincount(x, weights=None, minlength=0, length=None):
    n = int(length if length is not None else minlength)
    x = _raw(np.asarray(x))
    w = None if weights is None else _raw(np.asarray(weights))
    return _wrap(np.bincount(x, weights=w, minlength=n)[:n] if n else np.bincount(x, weights=w))


def _linspace(start, stop, num, **k):
    # jax 0.4.16 `_linspace`: iota/div then start*(1-step) + stop*step in the computation dtype,
    # endpoint concatenated
    num = int(num)
    start = np.float32(start)
    stop = np.float32(stop)
    if num == 1:
        return _wrap(np.array([start], np.float32))
    div = num - 1
    step = np.arange(div, dtype=np.float32) / np.float32(div)
    out = start * (np.float32(1) - step) + stop * step
    return _wrap(np.concatenate([out, np.array([stop], np.float32)]).astype(np.float32))


def _select(pred, a, b):
    return _wrap(np.where(_raw(np.asarray(pred)), _raw(np.asarray(a)), _raw(np.asarray(b))))


def _savez(*a, **k):
    return np.savez(*_tree_map(_raw, a), **_tree_map(_raw, k))


def _stack_tree(outs):
    first = outs[0]
    if isinstance(first, (tuple, list)):
        return type(first)(_stack_tree([o[i] for o in outs]) for i in range(len(first)))
    return _wrap(np.stack([np.asarray(_raw(o)) for o in outs], axis=0))


def vmap(f, in_axes=0, out_axes=0):
    assert in_axes == 0 and out_axes == 0, "shim: only in_axes=0 / out_axes=0"

    def g(*args):
        n = len(args[0])
        outs = [f(*[a[i] for a in args]) for i in range(n)]
        return _stack_tree(outs)

    return g


def jit(f=None, static_argnames=None, static_argnums=None, **k):
    if f is None:
        return lambda fn: jit(fn)
    f._clear_cache = lambda: None
    return f


def scan(f, init, xs, length=None):
    carry = init
    ys = []
    n = len(xs) if xs is not None else int(length)
    for i in range(n):
        x = _tree_map(lambda a: a[i], xs) if xs is not None else None
        carry, y = f(carry, x)
        ys.append(y)
    return carry, (_stack_tree(ys) if ys and ys[0] is not None else None)


class BCOO:
    """(data, indices (nnz, 2)) sparse matrix supporting ``M @ v`` only (cF:1355)."""

    def __init__(self, args, shape):
        data, indices = args
        self.data = np.asarray(_raw(data), dtype=np.float32)
        self.indices = np.asarray(_raw(indices))
        self.shape = tuple(int(s) for s in shape)

    def __matmul__(self, v):
        v = np.asarray(_raw(v), dtype=np.float32)
        out = np.zeros(self.shape[0], dtype=np.float32)
        np.add.at(out, self.indices[:, 0], self.data * v[self.indices[:, 1]])
        return _wrap(out)


def install():
    """Register the fake ``jax`` / ``pyevtk`` modules in sys.modules (idempotent)."""
    if "jax" in sys.modules and getattr(sys.modules["jax"], "_gomelt_shim", False):
        return sys.modules["jax"]
    jax = types.ModuleType("jax")
    jax._gomelt_shim = True
    jnp = types.ModuleType("jax.numpy")
    for name in ("matmul", "clip", "maximum", "minimum", "sqrt", "floor", "concatenate", "round", "exp",
                 "divmod", "stack", "tile", "repeat", "logical_and", "isclose", "diag", "sum", "diff", "abs",
                 "where", "multiply", "mean", "cumsum", "transpose", "dot", "power", "log", "argmax", "argmin",
                 "max", "min", "any", "all", "squeeze", "expand_dims", "take", "outer", "sign", "ceil", "mod"):
        setattr(jnp, name, _lift(getattr(np, name)))
    jnp.array = _array
    jnp.asarray = _array
    jnp.zeros = lambda shape, dtype=None: _wrap(np.zeros(_tree_map(_int, shape), dtype=dtype or np.float32), cast=False)
    jnp.ones = lambda shape, dtype=None: _wrap(np.ones(_tree_map(_int, shape), dtype=dtype or np.float32), cast=False)
    jnp.zeros_like = _lift(np.zeros_like)
    jnp.ones_like = _lift(np.ones_like)
    jnp.arange = _creator(np.arange)
    jnp.linspace = _linspace
    jnp.bincount = _bincount
    jnp.size = lambda a, axis=None: int(np.size(_raw(a), axis))
    jnp.pi = np.pi
    jnp.ndarray = np.ndarray
    jnp.float32, jnp.int32, jnp.bool_ = np.float32, np.int32, np.bool_
    jnp.savez = _savez
    linalg = types.ModuleType("jax.numpy.linalg")
    linalg.det = _lift(np.linalg.det)
    jnp.linalg = linalg
    lax = types.ModuleType("jax.lax")
    lax.scan = scan
    lax.select = _select
    experimental = types.ModuleType("jax.experimental")
    sparse = types.ModuleType("jax.experimental.sparse")
    sparse.BCOO = BCOO
    experimental.sparse = sparse
    config = types.SimpleNamespace(update=lambda *a, **k: None)
    jax.numpy, jax.lax, jax.experimental, jax.config = jnp, lax, experimental, config
    jax.jit, jax.vmap = jit, vmap
    jax.Array = np.ndarray
    sys.modules.update({"jax": jax, "jax.numpy": jnp, "jax.numpy.linalg": linalg, "jax.lax": lax,
                        "jax.experimental": experimental, "jax.experimental.sparse": sparse})
    pyevtk = types.ModuleType("pyevtk")
    hl = types.ModuleType("pyevtk.hl")
    hl.gridToVTK = lambda *a, **k: None
    pyevtk.hl = hl
    sys.modules.update({"pyevtk": pyevtk, "pyevtk.hl": hl})
    return jax


def _int(v):
    return int(v) if isinstance(v, (np.ndarray, np.generic)) and np.ndim(v) == 0 else v


def load_reference(path="/root/reference/go_melt"):
    """Import the reference's computeFunctions (unmodified source) on top of the shim."""
    install()
    if path not in sys.path:
        sys.path.insert(0, path)
    import importlib

    return importlib.import_module("computeFunctions")
