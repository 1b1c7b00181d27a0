"""A minimal numpy-backed stand-in for the parts of jax / flax.linen that the reference's *loss, RoPE and
param-group* modules touch — TEST INFRASTRUCTURE, used only by tests/golden/make_golden.py in the build container
(where /root/reference exists and the real jax/flax cannot be installed) to execute those reference files UNMODIFIED
and record golden vectors.  It is not a JAX implementation: only single-device, eager, float64 numpy semantics.
Third-party semantics restated here: nn.softmax / nn.log_softmax are the max-subtracted
forms; jax.lax.psum / pmean over one device are the identity; jax.lax.cond picks a branch eagerly.
"""
from __future__ import annotations

import sys
import types

import numpy as np


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        arr = self.arr

        class _Setter:
            def set(self, v):
                out = np.array(arr, copy=True).view(Arr)
                out[idx] = v
                return out
        return _Setter()


class Arr(np.ndarray):
    @property
    def at(self):
        return _At(self)

    def astype(self, dtype, *a, **k):
        if dtype is None:
            return self
        return np.ndarray.astype(self, dtype, *a, **k).view(Arr)


def _wrap(x):
    if isinstance(x, np.ndarray) and not isinstance(x, Arr):
        return x.view(Arr)
    if isinstance(x, (tuple, list)):
        return type(x)(_wrap(v) for v in x)
    return x


class _NumpyProxy(types.ModuleType):
    """jax.numpy: forwards to numpy, wraps results so `.at[...]` works."""

    def __init__(self, name, target):
        super().__init__(name)
        self._t = target

    def __getattr__(self, name):
        attr = getattr(self._t, name)
        if callable(attr) and not isinstance(attr, type):
            def f(*a, **k):
                if "dtype" in k and k["dtype"] is None:
                    k.pop("dtype")
                return _wrap(attr(*a, **k))
            return f
        if isinstance(attr, types.ModuleType):
            return _NumpyProxy(self.__name__ + "." + name, attr)
        return attr


def _fill_diagonal(a, val, inplace=True):
    out = np.array(a, copy=True)
    np.fill_diagonal(out, val)
    return out.view(Arr)


def _softmax(x, axis=-1):
    x = np.asarray(x)
    e = np.exp(x - x.max(axis=axis, keepdims=True))
    return (e / e.sum(axis=axis, keepdims=True)).view(Arr)


def _log_softmax(x, axis=-1):
    x = np.asarray(x)
    s = x - x.max(axis=axis, keepdims=True)
    return (s - np.log(np.exp(s).sum(axis=axis, keepdims=True))).view(Arr)


def _gelu(x, approximate=True):
    x = np.asarray(x)
    assert approximate
    return (0.5 * x * (1 + np.tanh(np.sqrt(2 / np.pi) * (x + 0.044715 * x ** 3)))).view(Arr)


class _Variable:
    def __init__(self, value):
        self.value = value


class Module:
    """flax.linen.Module stand-in: annotated class attributes become constructor fields; setup() runs eagerly."""

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)

    def __init__(self, *args, **kwargs):
        fields = []
        for klass in reversed(type(self).__mro__):
            fields += [n for n in getattr(klass, "__annotations__", {}) if n not in fields]
        for n, v in zip(fields, args):
            object.__setattr__(self, n, v)
        given = set(fields[: len(args)])
        for n in fields:
            if n in kwargs:
                object.__setattr__(self, n, kwargs[n])
            elif n not in given:
                if not hasattr(type(self), n):
                    raise TypeError(f"missing field {n}")
                object.__setattr__(self, n, getattr(type(self), n))   # instance attribute: plain functions stay unbound
        if hasattr(self, "setup"):
            self.setup()

    def variable(self, collection, name, init_fn, *a):
        return _Variable(_wrap(init_fn(*a)))


def _initializer(*a, **k):
    return lambda *aa, **kk: None


def install():
    """Put the stand-ins into sys.modules as `jax`, `jax.numpy`, `jax.lax`, `flax`, `flax.linen`, ..."""
    jnp = _NumpyProxy("jax.numpy", np)
    jnp.fill_diagonal = _fill_diagonal
    jnp.ndarray = np.ndarray
    jnp.float32, jnp.float16, jnp.bfloat16 = np.float32, np.float16, np.float32
    jnp.inf, jnp.nan, jnp.dtype = np.inf, np.nan, np.dtype

    jax = types.ModuleType("jax")
    jax.numpy = jnp
    jax.device_count = lambda: 1
    jax.local_device_count = lambda: 1
    lax = types.ModuleType("jax.lax")
    lax.psum = lambda x, axis_name=None: x
    lax.pmean = lambda x, axis_name=None: x
    lax.cond = lambda pred, t, f, operand=None: (t if pred else f)(operand)
    lax.axis_index = lambda name: 0
    jax.lax = lax
    jnn = types.ModuleType("jax.nn")
    jinit = types.ModuleType("jax.nn.initializers")
    jinit.truncated_normal = _initializer
    jnn.initializers = jinit
    jnn.gelu = _gelu
    jax.nn = jnn
    jrandom = types.ModuleType("jax.random")
    jax.random = jrandom
    jtu = types.ModuleType("jax.tree_util")

    def tree_map(fn, tree, *rest, is_leaf=None):
        if isinstance(tree, dict):
            return {k: tree_map(fn, v, *[r[k] for r in rest], is_leaf=is_leaf) for k, v in tree.items()}
        return fn(tree, *rest)
    jtu.tree_map = tree_map
    jax.tree_util = jtu
    jax.vmap = lambda f: (lambda x: np.stack([f(r) for r in x]))

    flax = types.ModuleType("flax")
    nn = types.ModuleType("flax.linen")
    nn.Module = Module
    nn.softmax, nn.log_softmax, nn.gelu = _softmax, _log_softmax, _gelu
    nn.compact = lambda f: f
    inits = types.ModuleType("flax.linen.initializers")
    for n in ("lecun_normal", "normal", "constant"):
        setattr(inits, n, _initializer)
    inits.zeros = _initializer()
    nn.initializers = inits
    for n in ("Dense", "Dropout", "LayerNorm", "Conv", "Sequential", "BatchNorm", "Partitioned", "silu"):
        setattr(nn, n, type(n, (), {}))
    flax.linen = nn
    tu = types.ModuleType("flax.traverse_util")

    def flatten_dict(d, sep=None, _pre=()):
        out = {}
        for k, v in d.items():
            if isinstance(v, dict):
                out.update(flatten_dict(v, None, _pre + (k,)))
            else:
                out[_pre + (k,)] = v
        if sep is not None and not _pre:
            out = {sep.join(k): v for k, v in out.items()}
        return out

    def unflatten_dict(d, sep=None):
        out = {}
        for k, v in d.items():
            parts = k.split(sep) if sep is not None else list(k)
            cur = out
            for p in parts[:-1]:
                cur = cur.setdefault(p, {})
            cur[parts[-1]] = v
        return out
    tu.flatten_dict, tu.unflatten_dict = flatten_dict, unflatten_dict
    flax.traverse_util = tu
    mods = {"jax": jax, "jax.numpy": jnp, "jax.lax": lax, "jax.nn": jnn, "jax.nn.initializers": jinit,
            "jax.random": jrandom, "jax.tree_util": jtu, "flax": flax, "flax.linen": nn,
            "flax.linen.initializers": inits, "flax.traverse_util": tu}
    sys.modules.update(mods)
    return mods
