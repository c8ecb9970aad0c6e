"""bf16-STORAGE model of the fp64 oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (same import rule as kfnet_oracle.py).

KFN_MODE_BF16 ("bf16 compute / fp32 accumulate", BASELINE configs[3..4]) differs from the reference's fp32 graph in
exactly one way: what is STORED is rounded to bfloat16 -- the packed weights, and every activation tensor that travels
between two conv layers.  Products are exact (8-bit x 8-bit significands) and sums are fp32.  The reference gives no
tolerance for such a mode ("no hard bound exists"), so instead of a free-hand bound the bf16 tests compare
the kernels with THIS model: the fp64 oracle with the same roundings injected at the same places,

    layer(x) = round_bf16( relu( conv_fp64( x, round_bf16(W) ) + b ) )        (network outputs stay unrounded)

How it is used.  A single LAYER fed identical bf16 operands agrees with exact arithmetic to one bf16 ulp (2^-8; asserted in
tests/test_gpu_conv_layers.py).  Across a 12- or 14-layer stack that does not extend to bit-level agreement: fp32-vs-fp64
accumulation noise (~1e-6) flips the occasional activation that sits on a bf16 rounding boundary, each flip perturbs every
downstream sum by a fraction of an ulp, which flips more activations -- two correct bf16 implementations are two independent
draws of the same rounding noise (measured on a B200: kernels vs this model 1.5e-2, either of them vs the exact oracle
2.6e-2, on SCoordNet's coordinate maps with the seeded synthetic weights).  So the model is a YARDSTICK, not a bit-exact
target: the bf16 tests (tests/test_gpu_networks.py, smoke(), bench.py's parity block) require the kernels' error against the
EXACT oracle to stay within a factor of the model's own error against the exact oracle (2 on a max-norm, 1.5 on a 99th
percentile): the kernels may lose nothing beyond what bf16 storage itself costs.

Roundings mirrored from the product (kfnet_b200/csrc/kfn_api.cu, conv_tc.cu, pointwise.cu):
  * first layers (SCoordNet conv1a, tower feat1): operand = u8 - 128 (exact), weights = bf16(fp32(w * 0.00625));
  * every other layer: weights bf16(w); biases fp32 (unrounded);
  * activations bf16 after bias + ReLU, except SCoordNet `prediction`, tower `feat7` (+ l2-normalise), U-Net
    `prediction` (logits) and FC `uncertainty`, which leave the kernels as fp32;
  * difference volume: bf16(f_cur - shifted f_prev) (diff_volume_kernel<bf16>).
"""
from __future__ import annotations

import contextlib

import numpy as np
import torch

from oracle import kfnet_oracle as O


def q(x: torch.Tensor) -> torch.Tensor:
    """round to nearest-even bfloat16, keep the dtype"""
    return x.to(torch.float32).to(torch.bfloat16).to(x.dtype)


def _Wq(weights, scope, name, dtype, first=False):
    w = torch.as_tensor(weights[f"{scope}/{name}/kernel"]).to(torch.float32)
    b = torch.as_tensor(weights[f"{scope}/{name}/bias"]).to(dtype)
    if first:   # 0.00625 folded in fp32 before the rounding (kfn_api.cu pack_op / pack_first_both)
        w = w * torch.tensor(0.00625, dtype=torch.float32)
    return q(w).to(dtype), b


def _stack(images, weights, scope, layers, dtype, round_last):
    x = torch.as_tensor(images).to(dtype) - 128.0           # exact in bf16; the preprocess scale lives in the weights
    n = len(layers)
    for i, (name, k, cout, s, relu) in enumerate(layers):
        w, b = _Wq(weights, scope, name, dtype, first=(i == 0))
        x = O.conv2d_same(x, w, b, s, relu)
        if i + 1 < n or round_last:
            x = q(x)
    return x


def scoordnet(images, weights, dtype=torch.float64):
    """cnn_wrapper/SCoordNet.py:18-44 with bf16 storage."""
    x = _stack(images, weights, "ScoreNet", O.SCOORDNET_LAYERS, dtype, round_last=False)
    return x[..., :3], torch.exp(x[..., 3:4])


def oflow_feat(images, weights, dtype=torch.float64):
    """KFNet/KFNet.py:315-341 with bf16 storage (feat7 + l2-normalise leave the kernel as fp32)."""
    return O.l2_normalize(_stack(images, weights, "Temporal", O.TOWER_LAYERS, dtype, round_last=False))


def oflownet(diff, weights, dtype=torch.float64, return_all=False):
    """cnn_wrapper/OFlowNet.py:17-56 with bf16 storage; same signature / return values as kfnet_oracle.oflownet."""
    cw = lambda n: _Wq(weights, "Temporal", n, dtype)
    c = lambda x, n, s=1: q(O.conv2d_same(x, *cw(n), s, True))
    d = lambda x, n: q(O.conv2d_transpose_same(x, *cw(n), True))
    diff = q(diff)
    conv0 = c(diff, "conv0")
    conv1a = c(conv0, "conv1a", 2)
    conv1b = c(conv1a, "conv1b")
    conv2a = c(conv1b, "conv2a", 2)
    conv2b = c(conv2a, "conv2b")
    conv3a = c(conv2b, "conv3a", 2)
    conv3b = c(conv3a, "conv3b")
    up2 = d(conv3b, "upconv2")
    conv4 = c(torch.cat([up2, conv2b], -1), "conv4")
    up1 = d(conv4, "upconv1")
    conv5 = c(torch.cat([up1, conv1b], -1), "conv5")
    up0 = d(conv5, "upconv0")
    conv6 = c(torch.cat([up0, conv0], -1), "conv6")
    pred = O.conv2d_same(conv6, *cw("prediction"), 1, False)
    logits = pred[..., 0].reshape(pred.shape[0], -1)
    prob = torch.softmax(logits, -1)
    feat = conv3b.reshape(conv3b.shape[0], -1)
    fc1 = q(O.dense(feat, *cw("fc1"), True))
    fc2 = q(O.dense(fc1, *cw("fc2"), True))
    unc = O.dense(fc2, *cw("uncertainty"), False)
    sigma_w = torch.exp(unc) * 1e-2
    if return_all:
        return prob, sigma_w, dict(logits=logits, conv3b=feat, conv0=conv0, conv6=conv6, log_sigma_w=unc,
                                   conv1b=conv1b, conv2b=conv2b, conv4=conv4, conv5=conv5, up0=up0)
    return prob, sigma_w


@contextlib.contextmanager
def _patched_oflownet():
    saved = O.oflownet
    O.oflownet = oflownet
    try:
        yield
    finally:
        O.oflownet = saved


def eval_sequence(images, weights, transform, seq_len=500, nis_gate=False, dtype=torch.float64):
    """kfnet_oracle.eval_sequence with every network under the bf16-storage model; the filter itself (softmax, flow,
    warp, Kalman update, transform) stays exact -- the product runs those stages in fp32 in every mode."""
    T = len(images)
    zc, zu, feats = [], [], []
    for t in range(T):
        c, u = scoordnet(images[t:t + 1], weights, dtype)
        zc.append(c); zu.append(u)
        feats.append(oflow_feat(images[t:t + 1], weights, dtype))
    with _patched_oflownet():
        return O.eval_sequence(np.asarray(images), weights, transform, seq_len=seq_len, nis_gate=nis_gate, dtype=dtype,
                               measure=(zc, zu), feats=feats)
