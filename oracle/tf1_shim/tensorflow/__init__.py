"""Minimal TensorFlow-1.x API shim on torch-CPU -- TEST INFRASTRUCTURE ONLY.

Purpose: TensorFlow 1.13 (environment.yml:89 of the reference) cannot be installed in this image,
but the reference's model-definition files (KFNet/KFNet.py, KFNet/util.py, cnn_wrapper/*.py,
tools/util.py) are py3-parseable.  Putting this package first on sys.path lets
tests/golden/make_golden.py IMPORT AND EXECUTE those files unmodified (eager instead of graph
mode), so the golden vectors under tests/golden/ pin the oracle's graph wiring (layer order,
slicing, concat order, offsets table, sampler index arithmetic, Kalman algebra) to the reference
source.  Only the ~55 tf.* symbols the eval path (incl. the printed loss / accuracy) touches exist; each follows the TF-1.13 op
definition.  Nothing under kfnet_b200/ imports this.
"""
import builtins as _b
import contextlib
import types

import numpy as np
import torch
import torch.nn.functional as F

_DTYPE = [torch.float64]          # working "float32" dtype (float64 for truth runs)
_VARS = {}                        # variable store: 'scope/layer/kernel' -> np.ndarray (TF layouts)
_SCOPE = []

AUTO_REUSE = "AUTO_REUSE"
float32 = "float32"
int32 = "int32"


def set_dtype(dt):
    _DTYPE[0] = dt


def set_variables(d):
    _VARS.clear()
    _VARS.update(d)


class _Shape(list):
    def as_list(self):
        return list(self)


torch.Tensor.get_shape = lambda self: _Shape(int(s) for s in self.shape)


def _dt(d):
    if d in ("float32", float32, torch.float32, torch.float64, np.float32):
        return _DTYPE[0]
    if d in ("int32", int32, torch.int32):
        return torch.int32
    raise ValueError(d)


def _t(x):
    if isinstance(x, torch.Tensor):
        return x
    a = np.asarray(x)
    if a.dtype.kind == "f":
        return torch.as_tensor(a).to(_DTYPE[0])
    return torch.as_tensor(a)


def _ints(shape):
    if isinstance(shape, torch.Tensor):
        return [int(v) for v in shape.reshape(-1).tolist()]
    return [int(v) if not isinstance(v, torch.Tensor) else int(v.item()) for v in shape]


@contextlib.contextmanager
def variable_scope(name, reuse=None):
    _SCOPE.append(name)
    try:
        yield
    finally:
        _SCOPE.pop()


@contextlib.contextmanager
def name_scope(name):
    yield


@contextlib.contextmanager
def device(name):
    yield


def _var(layer, what):
    return torch.as_tensor(_VARS["/".join(_SCOPE + [layer, what])]).to(_DTYPE[0])


# ---- elementwise / shape ops ---------------------------------------------------------------
def cast(x, dtype, name=None):
    return _t(x).to(_dt(dtype))


def constant(x, dtype=None):
    return _t(np.asarray(x))


def convert_to_tensor(x, dtype=None):
    return cast(x, dtype) if dtype is not None else _t(x)


def multiply(a, b, name=None): return _t(a) * b
def subtract(a, b, name=None): return _t(a) - b
def add(a, b, name=None): return _t(a) + b
def div(a, b, name=None): return _t(a) / b
def square(x, name=None): return x * x
def sqrt(x, name=None): return torch.sqrt(x)
def exp(x, name=None): return torch.exp(x)
def floor(x, name=None): return torch.floor(x)
def equal(a, b): return _t(a) == b


def maximum(a, b, name=None):
    return torch.maximum(_t(a), torch.as_tensor(b, dtype=_t(a).dtype))


def minimum(a, b, name=None):
    return torch.minimum(_t(a), torch.as_tensor(b, dtype=_t(a).dtype))


def log(x, name=None): return torch.log(x)


def reduce_sum(x, axis=None, keepdims=False, name=None):
    x = _t(x)
    return x.sum() if axis is None else x.sum(dim=axis, keepdim=keepdims)


def reduce_mean(x, axis=None, keepdims=False, name=None):
    x = _t(x)
    return x.mean() if axis is None else x.mean(dim=axis, keepdim=keepdims)


def abs(x, name=None): return torch.abs(x)                       # noqa: A001 (mirrors tf.abs)


def matrix_inverse(x, name=None): return torch.linalg.inv(x)


def count_nonzero(x, name=None):
    return (_t(x) != 0).sum()


def _resize_nearest_neighbor(images, size, align_corners=False):
    """tf.image.resize_nearest_neighbor, TF-1.13 default (align_corners=False, no half-pixel centres):
    src = min(floor(dst * in / out), in - 1)."""
    assert not align_corners
    n, h, w, c = images.shape
    oh, ow = int(size[0]), int(size[1])
    ys = torch.clamp(torch.floor(torch.arange(oh, dtype=torch.float64) * (h / oh)).long(), max=h - 1)
    xs = torch.clamp(torch.floor(torch.arange(ow, dtype=torch.float64) * (w / ow)).long(), max=w - 1)
    return images[:, ys][:, :, xs]


image = types.SimpleNamespace(resize_nearest_neighbor=_resize_nearest_neighbor)


def clip_by_value(x, lo, hi):
    return torch.minimum(torch.maximum(x, lo.to(x.dtype)), hi.to(x.dtype))


def add_n(xs):
    acc = xs[0]
    for x in xs[1:]:
        acc = acc + x
    return acc


def slice(x, begin, size, name=None):
    idx = tuple(_b.slice(b, None if s == -1 else b + s) for b, s in zip(begin, size))
    return x[idx]


def concat(values, axis, name=None):
    return torch.cat([_t(x) for x in values], dim=axis)


def stack(xs, axis=0, name=None):
    return torch.stack([_t(np.asarray(x)) if not isinstance(x, torch.Tensor) else x for x in xs], dim=axis)


def reshape(x, shape, name=None):
    return x.reshape(_ints(shape))


def tile(x, multiples, name=None):
    return x.repeat(*_ints(multiples))


def expand_dims(x, axis, name=None):
    return _t(x).unsqueeze(axis)


def squeeze(x, axis=None, name=None):
    return x.squeeze() if axis is None else x.squeeze(axis)


def transpose(x, perm):
    return x.permute(*perm)


def matmul(a, b):
    return a @ b


def split(x, sizes, axis):
    return torch.split(x, sizes, dim=axis)


def ones(shape, dtype=None):
    return torch.ones(_ints(shape), dtype=_DTYPE[0])


def zeros(shape, dtype="float32"):
    return torch.zeros(_ints(shape), dtype=_dt(dtype))


def range(n):
    return torch.arange(int(n), dtype=torch.int32)


def shape(x):
    return torch.tensor(list(x.shape), dtype=torch.int32)


def gather(params, indices):
    return _t(params)[indices.to(torch.int64)]


# ---- tf.layers -------------------------------------------------------------------------------
def _same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def _conv2d(inputs, filters, kernel_size, strides=1, activation=None, use_bias=True, padding="SAME",
            trainable=True, reuse=None, name=None):
    assert padding == "SAME"
    w = _var(name, "kernel")                       # HWIO
    assert tuple(w.shape) == (kernel_size, kernel_size, inputs.shape[-1], filters), (name, w.shape)
    x = inputs.permute(0, 3, 1, 2)
    pt, pb = _same_pad(x.shape[2], kernel_size, strides)
    pl, pr = _same_pad(x.shape[3], kernel_size, strides)
    x = F.pad(x, (pl, pr, pt, pb))
    y = F.conv2d(x, w.permute(3, 2, 0, 1), _var(name, "bias") if use_bias else None, stride=strides)
    y = y.permute(0, 2, 3, 1)
    return activation(y) if activation is not None else y


def _conv2d_transpose(inputs, filters, kernel_size, strides=1, activation=None, use_bias=True,
                      padding="SAME", trainable=True, reuse=None, name=None):
    # SAME conv2d_transpose == gradient of the SAME forward conv w.r.t. its input: computed here
    # literally as that gradient (autograd of F.conv2d), independent of the oracle's closed form.
    assert padding == "SAME"
    w = _var(name, "kernel")                       # [kh, kw, Cout, Cin]
    assert tuple(w.shape) == (kernel_size, kernel_size, filters, inputs.shape[-1]), (name, w.shape)
    n, h, wd, cin = inputs.shape
    oh, ow = h * strides, wd * strides
    big = torch.zeros(n, filters, oh, ow, dtype=inputs.dtype, requires_grad=True)
    pt, pb = _same_pad(oh, kernel_size, strides)
    pl, pr = _same_pad(ow, kernel_size, strides)
    # forward conv maps [.., filters(=its in-ch)] -> [.., cin(=its out-ch)] with HWIO kernel w
    fwd = F.conv2d(F.pad(big, (pl, pr, pt, pb)), w.permute(3, 2, 0, 1), None, stride=strides)
    (g,) = torch.autograd.grad(fwd, big, grad_outputs=inputs.permute(0, 3, 1, 2).contiguous())
    y = g.detach()
    if use_bias:
        y = y + _var(name, "bias").view(1, -1, 1, 1)
    y = y.permute(0, 2, 3, 1)
    return activation(y) if activation is not None else y


def _dense(inputs, units, activation=None, trainable=True, reuse=None, name=None):
    y = inputs @ _var(name, "kernel") + _var(name, "bias")
    return activation(y) if activation is not None else y


layers = types.SimpleNamespace(conv2d=_conv2d, conv2d_transpose=_conv2d_transpose, dense=_dense)


# ---- tf.nn -----------------------------------------------------------------------------------
def _l2_normalize(x, axis=-1, epsilon=1e-12):
    ss = (x * x).sum(dim=axis, keepdim=True)
    return x * torch.rsqrt(torch.clamp(ss, min=epsilon))


nn = types.SimpleNamespace(
    relu=torch.relu,
    tanh=torch.tanh,
    softmax=lambda x, axis=-1, name=None: torch.softmax(x, dim=axis),
    l2_normalize=_l2_normalize,
)


# ---- tf.contrib.image --------------------------------------------------------------------------
def _translate(images, translations, interpolation="NEAREST"):
    """tf.contrib.image.translate: translations_to_projective_transforms gives [1,0,-dx,0,1,-dy,0,0];
    transform() maps OUTPUT (x,y) -> INPUT (x - dx, y - dy), zero outside, NEAREST."""
    assert interpolation == "NEAREST"
    dx, dy = int(translations[0]), int(translations[1])
    n, h, w, c = images.shape
    out = torch.zeros_like(images)
    ys = torch.arange(h); xs = torch.arange(w)
    sy = ys - dy; sx = xs - dx
    vy = (sy >= 0) & (sy < h); vx = (sx >= 0) & (sx < w)
    if vy.any() and vx.any():
        out[:, ys[vy][:, None], xs[vx][None, :]] = images[:, sy[vy][:, None], sx[vx][None, :]]
    return out


contrib = types.SimpleNamespace(image=types.SimpleNamespace(translate=_translate),
                                layers=types.SimpleNamespace())
