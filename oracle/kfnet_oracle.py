"""CPU oracle for the KFNet hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (kfnet_b200/) never imports it and has no CPU fallback.

This is a CPU restatement (torch-CPU, fp64 or fp32) of the TF1 graph the reference builds for
`KFNet/eval.py`: SCoordNet conv stack, OFlowNet feature tower, feature-difference volume,
OFlowNet U-Net + FC head, soft-argmax flow, bilinear warp, process noise, Kalman update, NIS,
rigid transform and the sequential eval loop.  Every function cites the reference file:line it
follows (paths relative to /root/reference).

PARITY PINNING.  The reference ships no tests, golden vectors or checkpoints, and its arithmetic
lives in tensorflow-gpu==1.13.1 (environment.yml:89) which cannot be installed here (py3.12, no
network).  The oracle is therefore pinned two ways (see tests/test_oracle_*.py):
  1. op semantics (SAME padding side, conv2d_transpose crop, translate direction, bilinear edge
     weights, Kalman clamp) against naive numpy loops written from the TF-1.13 op definitions;
  2. graph wiring against golden vectors produced by EXECUTING the reference's own model files
     (KFNet/KFNet.py, cnn_wrapper/*.py, tools/util.py, KFNet/util.py) on a minimal TF1 API shim
     (oracle/tf1_shim) -- tests/golden/make_golden.py, fixtures committed under tests/golden/.
No real-TF output exists to compare against, so op-level semantics remain "parity unpinned" in
the strict sense; the graph-level wiring is pinned to the reference source.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# layer tables (cnn_wrapper/SCoordNet.py:21-32, KFNet/KFNet.py:318-338, cnn_wrapper/OFlowNet.py:19-41)
# (name, kernel, cout, stride, relu)
SCOORDNET_LAYERS = [
    ("conv1a", 3, 64, 1, True), ("conv1b", 3, 64, 1, True),
    ("conv2a", 3, 256, 2, True), ("conv2b", 3, 256, 1, True),
    ("conv3a", 3, 512, 2, True), ("conv3b", 3, 512, 1, True),
    ("conv4a", 3, 1024, 2, True), ("conv4b", 3, 1024, 1, True),
    ("conv5", 3, 512, 1, True), ("conv6", 3, 256, 1, True),
    ("conv7", 1, 128, 1, True), ("prediction", 1, 4, 1, False),
]
TOWER_LAYERS = [
    ("feat1", 3, 16, 1, True), ("feat2", 3, 32, 2, True), ("feat3", 3, 32, 1, True),
    ("feat4", 3, 64, 2, True), ("feat5", 3, 64, 1, True), ("feat6", 3, 128, 2, True),
    ("feat7", 3, 32, 1, False),
]
WINDOW = 8          # KFNet/KFNet.py:369  (64 // flow_sample_rate)
MIN_UNC = 1e-5      # KFNet/KFNet.py:71


# --------------------------------------------------------------------------------------------
# primitive ops with TF-1.13 semantics
def conv2d_same(x, w, b, stride=1, relu=True):
    """tf.layers.conv2d(padding='SAME') + bias + optional ReLU (cnn_wrapper/network.py:116-135).

    x NHWC, w HWIO, b [O].  SAME: out = ceil(in/s); pad_total = max((out-1)*s + k - in, 0);
    pad_before = pad_total // 2 (so 3x3/s2 on even input pads only bottom/right)."""
    kh, kw, cin, cout = w.shape
    n, h, wd, c = x.shape
    assert c == cin
    oh, ow = -(-h // stride), -(-wd // stride)
    ph = max((oh - 1) * stride + kh - h, 0)
    pw = max((ow - 1) * stride + kw - wd, 0)
    xt = x.permute(0, 3, 1, 2)
    xt = F.pad(xt, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
    y = F.conv2d(xt, w.permute(3, 2, 0, 1).contiguous(), b, stride=stride)   # contiguous: CPU autograd needs it
    if relu:
        y = torch.relu(y)
    return y.permute(0, 2, 3, 1).contiguous()


def conv2d_transpose_same(x, w, b, relu=True):
    """tf.layers.conv2d_transpose(k=3, s=2, padding='SAME') (cnn_wrapper/network.py:418-437).

    w is [kh, kw, Cout, Cin].  Output 2n; out[j] = sum_{i,k: 2i+k=j} in[i] w[k], j in [0,2n):
    the un-padded transposed conv (length 2n+1) cropped to its first 2n rows/cols (App. A.3)."""
    kh, kw, cout, cin = w.shape
    n, h, wd, c = x.shape
    assert c == cin and kh == 3 and kw == 3
    y = F.conv_transpose2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1).contiguous(), b, stride=2, padding=0)
    y = y[:, :, : 2 * h, : 2 * wd]
    if relu:
        y = torch.relu(y)
    return y.permute(0, 2, 3, 1).contiguous()


def dense(x, w, b, relu):
    """tf.layers.dense (cnn_wrapper/OFlowNet.py:50-55); w [in,out]."""
    y = x @ w + b
    return torch.relu(y) if relu else y


def preprocess(images):
    """(img - 128) * 0.00625 (cnn_wrapper/SCoordNet.py:34-37, KFNet/KFNet.py:317)."""
    return (images - 128.0) * 0.00625


def l2_normalize(x, eps=1e-12):
    """tf.nn.l2_normalize(axis=-1) = x * rsqrt(max(sum x^2, eps)) (KFNet/KFNet.py:340)."""
    ss = (x * x).sum(-1, keepdim=True)
    return x * torch.rsqrt(torch.clamp(ss, min=eps))


def translate_nearest(img, dx, dy):
    """tf.contrib.image.translate(img, [dx,dy], 'NEAREST'): out[y,x] = img[y-dy, x-dx], 0 outside (App. A.5)."""
    n, h, w, c = img.shape
    out = torch.zeros_like(img)
    ys0, ys1 = max(0, dy), min(h, h + dy)
    xs0, xs1 = max(0, dx), min(w, w + dx)
    if ys1 > ys0 and xs1 > xs0:
        out[:, ys0:ys1, xs0:xs1] = img[:, ys0 - dy:ys1 - dy, xs0 - dx:xs1 - dx]
    return out


def get_pixel_map(batch, height, width, dtype):
    """KFNet/util.py:42-63 (normalize=False): [...,0]=x, [...,1]=y integer grid."""
    ys, xs = torch.meshgrid(torch.arange(height, dtype=dtype), torch.arange(width, dtype=dtype), indexing="ij")
    return torch.stack([xs, ys], -1)[None].repeat(batch, 1, 1, 1)


def bilinear_sampler(imgs, coords, return_taps=False):
    """tools/util.py:3-94.  4-tap gather; weights from CLIPPED neighbours (:56-59); add_n order (:90-93)."""
    n, h, w, c = imgs.shape
    cx, cy = coords[..., 0:1], coords[..., 1:2]
    x0 = torch.floor(cx); x1 = x0 + 1
    y0 = torch.floor(cy); y1 = y0 + 1
    x0s = x0.clamp(0, w - 1); x1s = x1.clamp(0, w - 1)
    y0s = y0.clamp(0, h - 1); y1s = y1.clamp(0, h - 1)
    wx0 = x1s - cx; wx1 = cx - x0s
    wy0 = y1s - cy; wy1 = cy - y0s
    base = (torch.arange(n, dtype=imgs.dtype) * (h * w)).view(n, 1, 1, 1)
    flat = imgs.reshape(-1, c)

    def g(xs, ys):
        idx = (xs + base + ys * w).to(torch.int64).reshape(-1)
        return flat[idx].reshape(coords.shape[0], coords.shape[1], coords.shape[2], c)

    im00, im01, im10, im11 = g(x0s, y0s), g(x0s, y1s), g(x1s, y0s), g(x1s, y1s)
    w00, w01, w10, w11 = wx0 * wy0, wx0 * wy1, wx1 * wy0, wx1 * wy1
    out = ((w00 * im00 + w01 * im01) + w10 * im10) + w11 * im11
    if return_taps:
        return out, torch.cat([x0, y0], -1).to(torch.int32)
    return out


def apply_transform(coords, T):
    """KFNet/util.py:12-40: (T @ [x,y,z,1])[:3] per pixel; T 4x4."""
    ones = torch.ones_like(coords[..., :1])
    hom = torch.cat([coords, ones], -1)
    return (hom @ T.T.to(coords.dtype))[..., :3]


# --------------------------------------------------------------------------------------------
def _W(weights, scope, name, dtype):
    return (torch.as_tensor(weights[f"{scope}/{name}/kernel"]).to(dtype),
            torch.as_tensor(weights[f"{scope}/{name}/bias"]).to(dtype))


def scoordnet(images, weights, dtype=torch.float32, return_all=False):
    """cnn_wrapper/SCoordNet.py:18-44.  images [B,480,640,3] (u8 or float) -> coord [B,60,80,3], sigma [B,60,80,1]."""
    x = preprocess(torch.as_tensor(images).to(dtype))
    acts = {}
    for name, k, cout, s, relu in SCOORDNET_LAYERS:
        w, b = _W(weights, "ScoreNet", name, dtype)
        x = conv2d_same(x, w, b, s, relu)
        if return_all:
            acts[name] = x
    coord, unc = x[..., :3], torch.exp(x[..., 3:4])
    if return_all:
        return coord, unc, acts
    return coord, unc


def oflow_feat(images, weights, dtype=torch.float32):
    """KFNet/KFNet.py:315-341: 7-conv tower + l2_normalize -> [B,60,80,32]."""
    x = preprocess(torch.as_tensor(images).to(dtype))
    for name, k, cout, s, relu in TOWER_LAYERS:
        w, b = _W(weights, "Temporal", name, dtype)
        x = conv2d_same(x, w, b, s, relu)
    return l2_normalize(x)


def coord_volume(f1, f2, window=WINDOW):
    """KFNet/KFNet.py:343-359.  diff[...,k*C:(k+1)*C] = f2 - translate(f1, [-ox,-oy]), k = i*8+j, (ox,oy)=(j-4,i-4)."""
    half = window // 2
    diffs, offs = [], []
    for i in range(window):          # y
        for j in range(window):      # x
            ox, oy = j - half, i - half
            diffs.append(f2 - translate_nearest(f1, -ox, -oy))
            offs.append([ox, oy])
    return torch.cat(diffs, -1), torch.tensor(offs, dtype=f1.dtype)


def oflownet(diff, weights, dtype=torch.float32, return_all=False):
    """cnn_wrapper/OFlowNet.py:17-56.  diff [P,8,8,32] -> prob [P,64], sigma_w [P,1] (+ logits, bottleneck)."""
    cw = lambda n: _W(weights, "Temporal", n, dtype)
    conv0 = conv2d_same(diff, *cw("conv0"), 1, True)
    conv1a = conv2d_same(conv0, *cw("conv1a"), 2, True)
    conv1b = conv2d_same(conv1a, *cw("conv1b"), 1, True)
    conv2a = conv2d_same(conv1b, *cw("conv2a"), 2, True)
    conv2b = conv2d_same(conv2a, *cw("conv2b"), 1, True)
    conv3a = conv2d_same(conv2b, *cw("conv3a"), 2, True)
    conv3b = conv2d_same(conv3a, *cw("conv3b"), 1, True)
    up2 = conv2d_transpose_same(conv3b, *cw("upconv2"), True)
    conv4 = conv2d_same(torch.cat([up2, conv2b], -1), *cw("conv4"), 1, True)
    up1 = conv2d_transpose_same(conv4, *cw("upconv1"), True)
    conv5 = conv2d_same(torch.cat([up1, conv1b], -1), *cw("conv5"), 1, True)
    up0 = conv2d_transpose_same(conv5, *cw("upconv0"), True)
    conv6 = conv2d_same(torch.cat([up0, conv0], -1), *cw("conv6"), 1, True)
    pred = conv2d_same(conv6, *cw("prediction"), 1, False)          # [P,8,8,1]
    logits = pred[..., 0].reshape(pred.shape[0], -1)                 # [P,64] row-major (i,j)
    prob = torch.softmax(logits, -1)
    feat = conv3b.reshape(conv3b.shape[0], -1)                       # [P,128]
    fc1 = dense(feat, *cw("fc1"), True)
    fc2 = dense(fc1, *cw("fc2"), True)
    unc = dense(fc2, *cw("uncertainty"), False)
    sigma_w = torch.exp(unc) * 1e-2
    if return_all:
        return prob, sigma_w, dict(logits=logits, conv3b=feat, conv0=conv0, conv6=conv6, log_sigma_w=unc,
                                   conv1b=conv1b, conv2b=conv2b, conv4=conv4, conv5=conv5, up0=up0)
    return prob, sigma_w


def flow_from_prob(prob, offsets, batch, height, width):
    """KFNet/KFNet.py:381-386: flow = prob @ offsets; pixel_map = grid + flow."""
    flow = (prob @ offsets).reshape(batch, height, width, 2)
    return flow, get_pixel_map(batch, height, width, prob.dtype) + flow


def process_predict(pixel_map, sigma_w, coord1, unc1, return_taps=False):
    """KFNet/KFNet.py:388-401: warp the previous state and add process noise."""
    if return_taps:
        temp_coord, taps = bilinear_sampler(coord1, pixel_map, True)
    else:
        temp_coord = bilinear_sampler(coord1, pixel_map)
    last_unc = bilinear_sampler(unc1, pixel_map)
    last_var = torch.clamp(last_unc * last_unc, min=MIN_UNC * MIN_UNC)
    tv = sigma_w.reshape(unc1.shape)
    trans_var = torch.clamp(tv * tv, min=MIN_UNC * MIN_UNC)
    temp_unc = torch.sqrt(trans_var + last_var)
    if return_taps:
        return temp_coord, temp_unc, taps
    return temp_coord, temp_unc


def build_oflownet(f1, f2, coord1, unc1, weights, dtype=torch.float32, return_all=False):
    """KFNet/KFNet.py:361-403 (BuildOFlowNet)."""
    b, h, w, c = f1.shape
    diff, offsets = coord_volume(f1, f2, WINDOW)
    diff = diff.reshape(-1, WINDOW, WINDOW, c)
    prob, sigma_w, extra = oflownet(diff, weights, dtype, True)
    flow, pixel_map = flow_from_prob(prob, offsets, b, h, w)
    temp_coord, temp_unc, taps = process_predict(pixel_map, sigma_w, coord1, unc1, True)
    if return_all:
        extra.update(prob=prob, sigma_w=sigma_w, flow=flow, taps=taps, diff=diff,
                     argmax=prob.argmax(-1).to(torch.int32))
        return temp_coord, temp_unc, extra
    return temp_coord, temp_unc


def build_kf_coord(last_coord, last_unc, meas_coord, meas_unc):
    """KFNet/KFNet.py:148-162 (BuildKFCoord)."""
    lv = last_unc * last_unc
    mv = meas_unc * meas_unc
    K = lv / (lv + mv)
    omk = torch.clamp(1.0 - K, min=0.0)
    kf = omk * last_coord + K * meas_coord
    return kf, torch.sqrt(omk * lv)


def get_nis(meas_coord, meas_unc, temp_coord, temp_unc):
    """KFNet/KFNet.py:164-184: (z - x^-)^2 / (sqrt(P^- + R))^2 per xyz."""
    inno = meas_coord - temp_coord
    iu = torch.sqrt(temp_unc * temp_unc + meas_unc * meas_unc)
    return (inno * inno) / (iu * iu)


# --------------------------------------------------------------------------------------------
@dataclass
class FrameOut:
    measure_coord: torch.Tensor
    measure_unc: torch.Tensor
    temp_coord: torch.Tensor
    temp_unc: torch.Tensor
    kf_coord: torch.Tensor
    kf_unc: torch.Tensor
    nis: torch.Tensor
    out_map: torch.Tensor       # [60,80,4] = (T*x+, 1/sigma+) -- what eval.py np.save()s
    flow: torch.Tensor | None = None
    taps: torch.Tensor | None = None
    argmax: torch.Tensor | None = None


def eval_sequence(images, weights, transform, seq_len=500, nis_gate=False, dtype=torch.float32,
                  faithful_redundancy=False, measure=None, feats=None):
    """KFNet/eval.py:31-126 + KFNet/train.py:67-71,241-266, one test range [0, T).

    Pair stream: iteration 0 is (frame 1, frame 0) and is re-initialised (i % seq_len == 0), then
    (i-1, i).  The estimated frame is the second of the pair.  `transform` is the ALREADY INVERTED
    4x4 (train.py:51-52).  faithful_redundancy=True recomputes SCoordNet + tower on both frames of
    every pair exactly like the reference (batch of 2, KFNet.py:74-75); False computes each frame once
    (identical numbers).  Returns list[FrameOut] indexed by frame.
    """
    T = len(images)
    Tm = torch.as_tensor(np.asarray(transform)).to(dtype)
    outs = []
    if not faithful_redundancy:
        if measure is None:
            zc, zu = [], []
            for t in range(T):
                c, u = scoordnet(images[t:t + 1], weights, dtype)
                zc.append(c); zu.append(u)
        else:
            zc, zu = measure
        if feats is None:
            feats = [oflow_feat(images[t:t + 1], weights, dtype) for t in range(T)]
    last_coord = last_unc = None
    for i in range(T):
        prev = 1 if i == 0 else i - 1
        if faithful_redundancy:
            pair = torch.as_tensor(np.stack([images[prev], images[i]]))
            c2, u2 = scoordnet(pair, weights, dtype)
            f2 = oflow_feat(pair, weights, dtype)
            z, zunc, fa, fb = c2[1:2], u2[1:2], f2[0:1], f2[1:2]
        else:
            z, zunc, fa, fb = zc[i], zu[i], feats[prev], feats[i]
        if i % seq_len == 0:
            # eval.py:94-101: state and outputs := measurement (KF work of this iteration is discarded)
            if faithful_redundancy:   # the reference still executes the whole graph in this sess.run (timing fidelity)
                lc0 = last_coord if last_coord is not None else torch.zeros_like(z)
                lu0 = last_unc if last_unc is not None else torch.ones_like(zunc)
                build_oflownet(fa, fb, lc0, lu0, weights, dtype)
            fo = FrameOut(z, zunc, z, zunc, z, zunc, torch.zeros_like(z),
                          torch.cat([apply_transform(z, Tm), 1.0 / zunc], -1)[0])
            last_coord, last_unc = z, zunc
        else:
            tc, tu, ex = build_oflownet(fa, fb, last_coord, last_unc, weights, dtype, True)
            kc, ku = build_kf_coord(tc, tu, z, zunc)
            nis = get_nis(z, zunc, tc, tu)
            tkf = apply_transform(kc, Tm)
            if nis_gate:   # eval.py:87-92
                m = (nis.sum(-1, keepdim=True) > 7.815).to(dtype)
                tkf = m * apply_transform(z, Tm) + (1.0 - m) * tkf
            fo = FrameOut(z, zunc, tc, tu, kc, ku, nis, torch.cat([tkf, 1.0 / ku], -1)[0],
                          ex["flow"], ex["taps"], ex["argmax"])
            last_coord, last_unc = kc, ku
        outs.append(fo)
    return outs


def kf_coord_batch(images4, weights, dtype=torch.float32):
    """KFNet/KFNet.py:102-146 (GetKFCoordBatch): 4-frame unrolled filter, frame 1 = measurement."""
    c, u = scoordnet(images4, weights, dtype)
    f = oflow_feat(images4, weights, dtype)
    tcs, tus, kcs, kus = [c[0:1]], [u[0:1]], [c[0:1]], [u[0:1]]
    for t in range(1, 4):
        tc, tu = build_oflownet(f[t - 1:t], f[t:t + 1], kcs[-1], kus[-1], weights, dtype)
        kc, ku = build_kf_coord(tc, tu, c[t:t + 1], u[t:t + 1])
        tcs.append(tc); tus.append(tu); kcs.append(kc); kus.append(ku)
    return torch.cat(tcs), torch.cat(tus), torch.cat(kcs), torch.cat(kus)


# --------------------------------------------------------------------------------------------
# eval-side host metrics (KFNet/eval.py:10-29)
def nis_in_band_fraction(nis):
    a = np.asarray(nis).reshape(-1)
    a = a[a > 0.0]
    if a.size == 0:
        return 0.0
    return float(((a > 0.0157) & (a < 2.706)).sum()) / float(a.size)


def dist_error(coords, gt_coords, mask):
    d = np.sqrt(np.square(np.asarray(coords) - np.asarray(gt_coords)).sum(-1)) * np.asarray(mask)[..., 0]
    v = d.reshape(-1)
    v = v[v > 0]
    return (float(np.median(v)) * 100.0 if v.size else float("nan")), d * 100.0


def coord_loss_with_uncertainty(pred, unc, gt_full, mask_full, transform, dist_threshold=0.05, min_uncertainty=1e-5):
    """KFNet.CoordLossWithUncertainty as KF_fusion calls it (KFNet.py:192-232, train.py:254-259): `transform` applied
    to the prediction, full-resolution labels nearest-neighbour downsampled to the prediction grid (TF-1.13
    resize_nearest_neighbor: src = floor(dst * in / out)), mask kept where == 1.0 (`mask_full=None`: no mask, then
    valid_pixel = B*H*W without the +1).  Note the reference's `tf.minimum(loss_map, -2.0)` (sic, KFNet.py:217)."""
    pred = np.asarray(pred, np.float64); unc = np.asarray(unc, np.float64)
    T = np.asarray(transform, np.float64)
    pred = pred @ T[:3, :3].T + T[:3, 3]
    b, h, w, _ = pred.shape
    def down(a):
        a = np.asarray(a, np.float64)
        ys = np.minimum(np.floor(np.arange(h) * (a.shape[1] / h)).astype(int), a.shape[1] - 1)
        xs = np.minimum(np.floor(np.arange(w) * (a.shape[2] / w)).astype(int), a.shape[2] - 1)
        return a[:, ys][:, :, xs]
    gt = down(gt_full)
    d = np.square(pred - gt).sum(-1, keepdims=True)
    u = np.maximum(unc, min_uncertainty)
    loss_map = np.minimum(3.0 * np.log(u) + d / (2.0 * np.square(u)), -2.0)
    if mask_full is not None:
        mask = (down(mask_full) == 1.0).astype(np.float64)
        valid = mask.sum() + 1.0
        d = d * mask
        loss_map = loss_map * mask
    else:
        valid = float(b * h * w)
    loss = loss_map.sum() / valid
    acc = (valid - np.count_nonzero(np.maximum(d - dist_threshold * dist_threshold, 0))) / valid
    return float(loss), float(acc)

