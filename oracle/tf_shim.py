"""TEST INFRASTRUCTURE — a minimal `tensorflow` (1.x graph API) stand-in backed by torch, so that the UNMODIFIED reference
files `kpconv/kernels/convolution_ops.py` and `kpconv/models/network_blocks.py` can be imported and EXECUTED in a container
without TensorFlow (TensorFlow is not installable here and the reference ships no KPConv vectors).

Same idea as the `chainer` / `plyfile` stubs of oracle/ref_harness.py: the reference's own Python decides every shape, index,
formula and operation order; only the third-party primitives it calls (tf.gather, tf.matmul, tf.reduce_sum, ...: 45 functions,
all with textbook semantics, listed below with the TF 1.x documentation behaviour each one follows) are provided here, eagerly,
on torch tensors — which also makes every reference function differentiable, so reference GRADIENTS can be minted too.

    with tf_shim.installed(dtype=torch.float64):          # puts the stand-in into sys.modules['tensorflow']
        conv_ops = tf_shim.import_reference("kernels.convolution_ops")
        out = conv_ops.KPConv_ops(q, s, idx, f, K_points, K_values, extent, 'linear', 'sum')

`tf.float32` maps to the dtype given to `installed()` (float64 for golden vectors, float32 to mimic the reference's precision).
Variables created by the reference (`tf.Variable`) are recorded in creation order in `variables()`; `presets={scoped name:
value}` replaces the initial value of the variable with that name (e.g. 'conv2/weights', 'conv1/batch_normalization/gamma'),
so that a test can run the reference and the implementation under test on the very same weights.  Only `oracle/` scripts and `tests/` import this file.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types
from collections import namedtuple

import numpy as np
import torch

REFERENCE_KPCONV = "/root/reference/kpconv"

_state = {"dtype": torch.float64, "vars": [], "scope": [], "gen": None, "presets": {}}


class _Shape(list):
    def as_list(self):
        return list(self)


class T(torch.Tensor):
    """torch tensor whose `.shape` has TensorFlow's `as_list()` (convolution_ops.py:319)."""
    @property
    def shape(self):                                   # noqa: D401
        return _Shape(super().shape)


def wrap(x, dtype=None):
    """numpy / torch / python value -> shim tensor (floating values take the session dtype unless `dtype` is given)."""
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.as_tensor(np.asarray(x))
    if dtype is not None:
        t = t.to(_dt(dtype))
    elif t.is_floating_point():
        t = t.to(_state["dtype"])
    return t.as_subclass(T)


class _DType:
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return "tf." + self.name


float32, float64, int32, int64, bool_ = _DType("float32"), _DType("float64"), _DType("int32"), _DType("int64"), _DType("bool")


def _dt(d):
    if d is None:
        return None
    if isinstance(d, torch.dtype):
        return d
    if isinstance(d, _DType):
        return {"float32": _state["dtype"], "float64": torch.float64, "int32": torch.int32, "int64": torch.int64, "bool": torch.bool}[d.name]
    return torch.as_tensor(np.zeros(0, d)).dtype


def _ax(kw, axis=None):
    for k in ("axis", "reduction_indices"):
        if kw.get(k) is not None:
            return kw[k]
    return axis


def _keep(kw):
    return bool(kw.get("keep_dims", kw.get("keepdims", False)))


def _build() -> types.ModuleType:
    tf = types.ModuleType("tensorflow")
    tf.__doc__ = "torch-backed stand-in for the TensorFlow 1.x functions the KPConv reference calls (oracle/tf_shim.py)"
    tf.float32, tf.float64, tf.int32, tf.int64, tf.bool = float32, float64, int32, int64, bool_
    tf.newaxis = None

    @contextlib.contextmanager
    def variable_scope(name=None, *a, **k):
        _state["scope"].append(str(name))
        try:
            yield
        finally:
            _state["scope"].pop()
    tf.variable_scope = variable_scope
    tf.name_scope = variable_scope

    def Variable(initial_value, name=None, trainable=True, dtype=None, **_):
        v = wrap(initial_value, dtype).detach().clone().as_subclass(T)
        full = "/".join(_state["scope"] + [name or "Variable"])
        if full in _state["presets"]:                  # the harness supplies this variable's value (same shape required)
            pv = wrap(_state["presets"][full], dtype).detach().clone().as_subclass(T)
            if list(pv.shape) != list(v.shape):
                raise ValueError("preset for %s has shape %s, the reference creates %s" % (full, list(pv.shape), list(v.shape)))
            v = pv
        if trainable and v.is_floating_point():
            v.requires_grad_(True)
        _state["vars"].append((full, v))
        return v
    tf.Variable = Variable

    def constant(value, dtype=None, shape=None, name=None):
        t = wrap(value, dtype)
        if shape is not None:
            shp = tuple(int(v) for v in shape)
            t = (torch.full(shp, t.item(), dtype=t.dtype) if t.dim() == 0 else t.reshape(shp)).as_subclass(T)
        return t
    tf.constant = constant
    tf.convert_to_tensor = lambda v, dtype=None, name=None: wrap(v, dtype)
    tf.zeros = lambda shape, dtype=float32, name=None: torch.zeros(*_sh(shape), dtype=_dt(dtype)).as_subclass(T)
    tf.ones = lambda shape, dtype=float32, name=None: torch.ones(*_sh(shape), dtype=_dt(dtype)).as_subclass(T)
    tf.zeros_like = lambda x, dtype=None, name=None: torch.zeros_like(x, dtype=_dt(dtype))
    tf.ones_like = lambda x, dtype=None, name=None: torch.ones_like(x, dtype=_dt(dtype))
    tf.range = lambda *a, **k: torch.arange(*[int(v) for v in a]).to(_dt(k.get("dtype", int32))).as_subclass(T)

    def _sh(shape):
        if isinstance(shape, (int, np.integer)):
            return (int(shape),)
        return tuple(int(s) for s in shape)

    # ---- shape manipulation
    tf.shape = lambda x, name=None, out_type=int32: torch.as_tensor(list(torch.Tensor.size(x)), dtype=_dt(out_type)).as_subclass(T)
    tf.expand_dims = lambda x, axis=None, name=None, dim=None: torch.unsqueeze(x, axis if axis is not None else dim)
    tf.squeeze = lambda x, axis=None, name=None, squeeze_dims=None: (torch.squeeze(x) if (axis is None and squeeze_dims is None)
                                                                    else torch.squeeze(x, axis if axis is not None else squeeze_dims))
    tf.reshape = lambda x, shape, name=None: torch.reshape(x, tuple(int(s) for s in shape))
    tf.transpose = lambda x, perm=None, name=None: (x.permute(*perm) if perm is not None else x.permute(*reversed(range(x.dim()))))
    tf.tile = lambda x, multiples, name=None: x.repeat(*[int(m) for m in multiples])
    tf.concat = lambda values, axis, name=None: torch.cat([wrap(v) if not isinstance(v, torch.Tensor) else v for v in values], dim=axis)
    tf.stack = lambda values, axis=0, name=None: torch.stack(list(values), dim=axis)
    tf.cast = lambda x, dtype, name=None: wrap(x, dtype) if not isinstance(x, torch.Tensor) else x.to(_dt(dtype))

    # ---- gathers (tf.gather: rows of `params` along `axis` for an index tensor of any rank; batch_gather: per-row gather
    # along axis 1 with the leading dimension as batch; gather_nd: the last index dimension addresses leading dims)
    def gather(params, indices, validate_indices=None, name=None, axis=0):
        idx = torch.as_tensor(indices).long()
        if axis == 0:
            return params[idx]
        return torch.index_select(params, axis, idx.reshape(-1)).reshape(*params.shape[:axis], *idx.shape, *params.shape[axis + 1:])
    tf.gather = gather

    def batch_gather(params, indices, name=None):
        idx = indices.long()
        extra = params.dim() - idx.dim()
        ix = idx.reshape(*idx.shape, *([1] * extra)).expand(*idx.shape, *params.shape[idx.dim():])
        return torch.gather(params, idx.dim() - 1, ix)
    tf.batch_gather = batch_gather

    def gather_nd(params, indices, name=None):
        idx = indices.long()
        return params[tuple(idx[..., i] for i in range(idx.shape[-1]))]
    tf.gather_nd = gather_nd

    # ---- arithmetic
    tf.matmul = lambda a, b, transpose_a=False, transpose_b=False, name=None: torch.matmul(
        a.transpose(-1, -2) if transpose_a else a, b.transpose(-1, -2) if transpose_b else b)
    tf.add = lambda a, b, name=None: a + b
    tf.square = lambda x, name=None: x * x
    tf.sqrt = lambda x, name=None: torch.sqrt(x)
    tf.exp = lambda x, name=None: torch.exp(x)
    tf.sigmoid = lambda x, name=None: torch.sigmoid(x)
    tf.round = lambda x, name=None: torch.round(x)            # both round half to even
    tf.maximum = lambda a, b, name=None: torch.clamp(a, min=b) if not isinstance(b, torch.Tensor) else torch.maximum(a, b)
    tf.less = lambda a, b, name=None: a < b

    def _reduce(fn_all, fn_axis):
        def f(x, axis=None, keepdims=None, name=None, **kw):
            ax = _ax(kw, axis)
            kd = bool(keepdims) or _keep(kw)
            if ax is None:
                return fn_all(x)
            return fn_axis(x, ax, kd)
        return f
    tf.reduce_sum = _reduce(lambda x: x.sum(), lambda x, a, k: x.sum(dim=a, keepdim=k))
    tf.reduce_mean = _reduce(lambda x: x.mean(), lambda x, a, k: x.mean(dim=a, keepdim=k))
    # amax / amin split the gradient equally between tied entries, as tf.reduce_max / reduce_min do
    tf.reduce_max = _reduce(lambda x: x.max(), lambda x, a, k: torch.amax(x, dim=a, keepdim=k))
    tf.reduce_min = _reduce(lambda x: x.min(), lambda x, a, k: torch.amin(x, dim=a, keepdim=k))
    tf.reduce_any = _reduce(lambda x: x.any(), lambda x, a, k: x.any(dim=a, keepdim=k))
    tf.argmin = lambda x, axis=None, name=None, dimension=None, output_type=int64: torch.argmin(
        x, dim=axis if axis is not None else dimension).to(_dt(output_type))

    def one_hot(indices, depth, on_value=None, off_value=None, axis=None, dtype=None, name=None):
        oh = torch.nn.functional.one_hot(indices.long(), int(depth)).to(_dt(dtype or float32))
        if axis is not None and axis != -1:
            oh = oh.movedim(-1, axis)
        return oh.as_subclass(T)
    tf.one_hot = one_hot

    TopK = namedtuple("TopKV2", ["values", "indices"])
    tf.math = types.ModuleType("tensorflow.math")

    def top_k(x, k=1, sorted=True, name=None):
        # tf.math.top_k: largest first; equal values -> the lower index first.  A stable descending sort gives exactly that.
        k = int(k)
        order = torch.sort(x, dim=-1, descending=True, stable=True)[1][..., :k]
        return TopK(torch.gather(x, -1, order), order.to(torch.int32))
    tf.math.top_k = top_k
    tf.nn = types.ModuleType("tensorflow.nn")
    tf.nn.top_k = top_k
    tf.nn.leaky_relu = lambda features, alpha=0.2, name=None: torch.nn.functional.leaky_relu(features, alpha)
    tf.nn.relu = lambda x, name=None: torch.relu(x)
    tf.nn.dropout = lambda x, keep_prob=None, **_: x          # tests run the blocks with dropout disabled (keep_prob = 1)
    tf.nn.sparse_softmax_cross_entropy_with_logits = lambda labels=None, logits=None, name=None: torch.nn.functional.cross_entropy(
        logits, labels.long(), reduction="none")

    # ---- layers: tf.layers.batch_normalization(training=True) normalises with the batch mean and the BIASED batch variance
    # over all axes but the last, y = gamma * (x - mean) / sqrt(var + eps) + beta with fresh gamma = 1 / beta = 0 variables
    tf.layers = types.ModuleType("tensorflow.layers")

    def batch_normalization(x, axis=-1, momentum=0.99, epsilon=1e-3, center=True, scale=True, training=False, name=None, **_):
        c = int(x.shape[-1])
        with variable_scope(name or "batch_normalization"):
            gamma = Variable(torch.ones(c, dtype=_state["dtype"]), name="gamma")
            beta = Variable(torch.zeros(c, dtype=_state["dtype"]), name="beta")
            mm = Variable(torch.zeros(c, dtype=_state["dtype"]), name="moving_mean", trainable=False)
            mv = Variable(torch.ones(c, dtype=_state["dtype"]), name="moving_variance", trainable=False)
        if not (training is True or (isinstance(training, torch.Tensor) and bool(training))):
            return (x - mm) / torch.sqrt(mv + epsilon) * gamma + beta
        red = tuple(range(x.dim() - 1))
        mean = x.mean(dim=red, keepdim=True)
        var = ((x - mean) ** 2).mean(dim=red, keepdim=True)
        return (x - mean) / torch.sqrt(var + epsilon) * gamma + beta
    tf.layers.batch_normalization = batch_normalization

    # ---- random: tf.truncated_normal = normal re-drawn until |z| <= 2 sigma (here from the session generator)
    def truncated_normal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None, name=None):
        g = _state["gen"]
        shp = _sh(shape)
        z = torch.randn(*shp, generator=g, dtype=torch.float64)
        bad = z.abs() > 2
        while bool(bad.any()):
            z = torch.where(bad, torch.randn(*shp, generator=g, dtype=torch.float64), z)
            bad = z.abs() > 2
        return (z * stddev + mean).to(_dt(dtype)).as_subclass(T)
    tf.truncated_normal = truncated_normal
    tf.set_random_seed = lambda s: _state["gen"].manual_seed(int(s))
    tf.Tensor = torch.Tensor
    return tf


@contextlib.contextmanager
def installed(dtype=torch.float64, seed=0, presets=None):
    """sys.modules['tensorflow'] = the stand-in (plus `matplotlib` stubs: kernels/kernel_points.py imports pyplot at module
    level for a plotting helper that is never called) for the duration of the block; the reference modules imported inside are
    dropped again on exit so that nothing leaks into other tests."""
    _state.update(dtype=dtype, vars=[], scope=[], gen=torch.Generator().manual_seed(seed), presets=dict(presets or {}))
    saved = {k: sys.modules.get(k) for k in ("tensorflow", "matplotlib", "matplotlib.pyplot")}
    sys.modules["tensorflow"] = _build()
    if saved["matplotlib"] is None:
        mpl = types.ModuleType("matplotlib"); plt = types.ModuleType("matplotlib.pyplot"); mpl.pyplot = plt
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    before = set(sys.modules)
    path_added = REFERENCE_KPCONV not in sys.path
    if path_added:
        sys.path.insert(0, REFERENCE_KPCONV)
    try:
        yield sys.modules["tensorflow"]
    finally:
        for k in set(sys.modules) - before:
            f = getattr(sys.modules[k], "__file__", "") or ""
            if f.startswith(REFERENCE_KPCONV):
                del sys.modules[k]
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        if path_added:
            sys.path.remove(REFERENCE_KPCONV)


def import_reference(name: str):
    """Import `name` (e.g. 'kernels.convolution_ops') from the read-only reference tree; call inside `installed()`."""
    if not os.path.isdir(REFERENCE_KPCONV):
        raise FileNotFoundError(REFERENCE_KPCONV + " is not present (the reference only exists in the build container)")
    return importlib.import_module(name)


@contextlib.contextmanager
def presets(values: dict):
    """Inside the block, a `tf.Variable` whose scoped name is a key of `values` starts from that value."""
    old = _state["presets"]
    _state["presets"] = dict(old, **values)
    try:
        yield
    finally:
        _state["presets"] = old


def variables():
    """[(scoped name, tensor)] in creation order since `installed()` was entered."""
    return list(_state["vars"])


def reset_variables():
    _state["vars"] = []
