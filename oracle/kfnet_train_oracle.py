"""Training-side oracle of KFNet -- TEST INFRASTRUCTURE ONLY, same rules as kfnet_oracle.py.

Restates, in differentiable torch-CPU code, what the reference's training graph computes on top of the inference path:
the three losses of KFNet/train.py:287-296 (MeasureCoordLoss / TemporalCoordLoss / KFCoordLoss = uncertainty-weighted
coordinate loss + 50 x image-weighted smoothness, KFNet.py:192-298, 430-467), their 0.2 / 0.2 / 0.6 mix, the L2
regulariser over ALL trainable variables and one AdamOptimizer step with exponential learning-rate decay
(train.py:300-315).  Gradients come from torch autograd through kfnet_oracle's forward functions, so this file is the
truth a future CUDA backward (dgrad / wgrad of the conv stacks, backward of warp / soft-argmax / Kalman update) will be
compared with.  The product has no training path yet; nothing under kfnet_b200/ imports this.

Parity pinning: the loss FORWARD values are pinned to the reference's own methods executed on oracle/tf1_shim
(tests/golden/make_golden.py::training_losses -> kfnet_reference_training_on_shim.npz).  The optimiser step follows
the TF-1.13 definitions of tf.train.AdamOptimizer / exponential_decay / l2_regularizer (documented formulas; no
TensorFlow output exists offline: unpinned, like the op-level semantics of kfnet_oracle.py).
"""
import math

import numpy as np
import torch

from . import kfnet_oracle as O

MIN_UNCERTAINTY = 1e-5          # KFNet.py:71


def coord_loss_with_uncertainty(pred, unc, gt, mask, dist_threshold=0.05):
    """KFNet.CoordLossWithUncertainty as the TRAINING graph calls it (KFNet.py:244: no transform, no downsample; labels
    already on the 60x80 grid, mask in {0,1}).  loss_map = min(3 log s + |x - gt|^2 / (2 s^2), -2)  (sic: tf.minimum,
    KFNet.py:217 -- only pixels whose unclamped loss is below -2 carry gradient)."""
    d = ((pred - gt) ** 2).sum(-1, keepdim=True)
    u = torch.clamp(unc, min=MIN_UNCERTAINTY)
    loss_map = torch.clamp(3.0 * torch.log(u) + d / (2.0 * u * u), max=-2.0)
    valid = mask.sum() + 1.0
    loss = (mask * loss_map).sum() / valid
    acc = (valid - torch.count_nonzero(torch.clamp(mask * d - dist_threshold ** 2, min=0))) / valid
    return loss, acc


def smooth_loss(coord_map, img, mask=None):
    """KFNet.SmoothLoss (KFNet.py:430-467): squared coordinate gradients weighted by exp(-0.625 |image gradient|)."""
    gx = coord_map[:, :, :-1] - coord_map[:, :, 1:]
    gy = coord_map[:, :-1] - coord_map[:, 1:]
    ix = img[:, :, :-1] - img[:, :, 1:]
    iy = img[:, :-1] - img[:, 1:]
    sx = (gx ** 2).mean(-1, keepdim=True) * torch.exp(-0.625 * ix.abs().mean(-1, keepdim=True))
    sy = (gy ** 2).mean(-1, keepdim=True) * torch.exp(-0.625 * iy.abs().mean(-1, keepdim=True))
    if mask is None:
        return sx.mean() + sy.mean()
    return ((sx * mask[:, :, :-1]).sum() + (sy * mask[:, :-1]).sum()) / (mask.sum() + 1.0)


def reprojection_loss(coord_map, pixel_map, poses, mask=None):
    """KFNet.ReprojectionLoss (KFNet.py:404-428); not part of train.py's loss (pixel_map / poses are never passed)."""
    hom = torch.cat([coord_map, torch.ones_like(coord_map[..., :1])], -1)
    P = poses if poses.dim() == 3 else poses[None].expand(coord_map.shape[0], 4, 4)
    cam = torch.einsum("bij,bhwj->bhwi", P.to(coord_map.dtype), hom)[..., :3]     # ApplyTransform, KFNet/util.py:12-40
    ones = torch.ones_like(pixel_map[..., :1])
    a = O.l2_normalize(cam)
    b = O.l2_normalize(torch.cat([pixel_map, ones], -1))
    err = (a - b) ** 2
    if mask is not None:
        return (mask * err).sum() / (mask.sum() + 1.0 + 1.0)
    n, h, w, _ = coord_map.shape
    return err.sum() / (n * h * w + 1)


def training_losses(images4, gt8, mask8, weights, dtype=torch.float64):
    """train.py:268-298 `run`: images [4,H,W,3] (0..255), gt8 [4,H/8,W/8,3] already moved into the network's frame
    (ApplyTransform(gt, transform, inverse=True), train.py:280), mask8 [4,H/8,W/8,1] in {0,1}.
    Returns the dict of scalars the reference's `run` returns plus the mixed loss."""
    images4 = torch.as_tensor(images4).to(dtype)
    gt8 = torch.as_tensor(gt8).to(dtype); mask8 = torch.as_tensor(mask8).to(dtype)
    h8, w8 = gt8.shape[1:3]
    ys = torch.arange(h8) * (images4.shape[1] // h8); xs = torch.arange(w8) * (images4.shape[2] // w8)
    resize_images = images4[:, ys][:, :, xs]                      # tf.image.resize_nearest_neighbor (train.py:274)
    mc, mu = O.scoordnet(images4, weights, dtype)
    tc, tu, kc, ku = O.kf_coord_batch(images4, weights, dtype)
    out = {}
    for name, (c, u) in (("measure", (mc, mu)), ("temporal", (tc, tu)), ("kf", (kc, ku))):
        loss, acc = coord_loss_with_uncertainty(c, u, gt8, mask8)
        out[name + "_loss"] = loss + 50.0 * smooth_loss(c, resize_images, mask8)      # KFNet.py:250-252, 276-278, 295-297
        out[name + "_accuracy"] = acc
    out["loss"] = 0.2 * out["measure_loss"] + 0.2 * out["temporal_loss"] + 0.6 * out["kf_loss"]   # train.py:294-296
    return out


def l2_regularization(weights, scale):
    """tf.contrib.layers.apply_regularization(l2_regularizer(scale), TRAINABLE_VARIABLES) (train.py:301-303):
    scale * sum(w^2) / 2 over every kernel AND bias."""
    return scale * sum((w ** 2).sum() for w in weights.values()) / 2.0


def exponential_decay(base_lr, step, decay_steps, gamma):
    """tf.train.exponential_decay, staircase=False (train.py:306-310)."""
    return base_lr * gamma ** (step / decay_steps)


def adam_step(params, grads, state, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer (TF-1.13 docs): t += 1; lr_t = lr sqrt(1 - b2^t) / (1 - b1^t);
    m = b1 m + (1 - b1) g; v = b2 v + (1 - b2) g^2; p -= lr_t m / (sqrt(v) + eps)   ("epsilon hat" form)."""
    state["t"] = state.get("t", 0) + 1
    t = state["t"]
    lr_t = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    new = {}
    for k, p in params.items():
        g = grads[k]
        m = beta1 * state.get(("m", k), torch.zeros_like(p)) + (1 - beta1) * g
        v = beta2 * state.get(("v", k), torch.zeros_like(p)) + (1 - beta2) * g * g
        state[("m", k)], state[("v", k)] = m, v
        new[k] = p - lr_t * m / (v.sqrt() + eps)
    return new


def train_step(weights, images4, gt8, mask8, state, base_lr=1e-4, weight_decay=1e-4, decay_steps=80000, gamma=0.5,
               dtype=torch.float64):
    """One iteration of train.py's loop (flag defaults of train.py:25-37; stepvalue is per scene, train.py:329-345):
    loss + regulariser -> gradients of every variable -> Adam update.
    `weights`: {name: ndarray/tensor}; returns (new weights as tensors, dict of scalars, gradients)."""
    params = {k: torch.as_tensor(np.asarray(v) if not isinstance(v, torch.Tensor) else v).to(dtype).clone().requires_grad_(True)
              for k, v in weights.items()}
    out = training_losses(images4, gt8, mask8, params, dtype)
    total = out["loss"] + l2_regularization(params, weight_decay)
    grads = dict(zip(params, torch.autograd.grad(total, list(params.values()), allow_unused=True)))
    grads = {k: (g if g is not None else torch.zeros_like(params[k])) for k, g in grads.items()}
    lr = exponential_decay(base_lr, state.get("t", 0), decay_steps, gamma)
    with torch.no_grad():
        new = adam_step({k: p.detach() for k, p in params.items()}, grads, state, lr)
    return new, {k: float(v.detach()) for k, v in out.items()}, grads
