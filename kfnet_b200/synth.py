"""Seeded synthetic weights / image sequences / labels / transforms.

No checkpoints or datasets ship with the reference (README.md:78,110 are Google-Drive links), so
every parity test and the benchmark run on these.  Weight names and layouts are the TF variable
names of the reference graph:
  ScoreNet/<layer>/{kernel[kh,kw,Cin,Cout], bias[Cout]}        cnn_wrapper/SCoordNet.py:21-32
  Temporal/feat1..7, conv0..conv6, prediction /{kernel,bias}    KFNet/KFNet.py:318-338, cnn_wrapper/OFlowNet.py:19-41
  Temporal/upconv{2,1,0}/{kernel[kh,kw,Cout,Cin], bias}         cnn_wrapper/network.py:418-437
  Temporal/{fc1,fc2,uncertainty}/{kernel[in,out], bias}         cnn_wrapper/OFlowNet.py:50-55
"""
from __future__ import annotations

import numpy as np

# (scope, name, kind, k, cin, cout)   kind in {conv, deconv, dense}
LAYER_SPECS = [
    ("ScoreNet", "conv1a", "conv", 3, 3, 64), ("ScoreNet", "conv1b", "conv", 3, 64, 64),
    ("ScoreNet", "conv2a", "conv", 3, 64, 256), ("ScoreNet", "conv2b", "conv", 3, 256, 256),
    ("ScoreNet", "conv3a", "conv", 3, 256, 512), ("ScoreNet", "conv3b", "conv", 3, 512, 512),
    ("ScoreNet", "conv4a", "conv", 3, 512, 1024), ("ScoreNet", "conv4b", "conv", 3, 1024, 1024),
    ("ScoreNet", "conv5", "conv", 3, 1024, 512), ("ScoreNet", "conv6", "conv", 3, 512, 256),
    ("ScoreNet", "conv7", "conv", 1, 256, 128), ("ScoreNet", "prediction", "conv", 1, 128, 4),
    ("Temporal", "feat1", "conv", 3, 3, 16), ("Temporal", "feat2", "conv", 3, 16, 32),
    ("Temporal", "feat3", "conv", 3, 32, 32), ("Temporal", "feat4", "conv", 3, 32, 64),
    ("Temporal", "feat5", "conv", 3, 64, 64), ("Temporal", "feat6", "conv", 3, 64, 128),
    ("Temporal", "feat7", "conv", 3, 128, 32),
    ("Temporal", "conv0", "conv", 3, 32, 32), ("Temporal", "conv1a", "conv", 3, 32, 32),
    ("Temporal", "conv1b", "conv", 3, 32, 32), ("Temporal", "conv2a", "conv", 3, 32, 64),
    ("Temporal", "conv2b", "conv", 3, 64, 64), ("Temporal", "conv3a", "conv", 3, 64, 128),
    ("Temporal", "conv3b", "conv", 3, 128, 128), ("Temporal", "upconv2", "deconv", 3, 128, 64),
    ("Temporal", "conv4", "conv", 3, 128, 64), ("Temporal", "upconv1", "deconv", 3, 64, 32),
    ("Temporal", "conv5", "conv", 3, 64, 32), ("Temporal", "upconv0", "deconv", 3, 32, 16),
    ("Temporal", "conv6", "conv", 3, 48, 16), ("Temporal", "prediction", "conv", 3, 16, 1),
    ("Temporal", "fc1", "dense", 1, 128, 64), ("Temporal", "fc2", "dense", 1, 64, 32),
    ("Temporal", "uncertainty", "dense", 1, 32, 1),
]


def weight_shapes():
    """name -> shape for every variable the eval graph restores (scopes 'ScoreNet' + 'Temporal', eval.py:66-68)."""
    out = {}
    for scope, name, kind, k, cin, cout in LAYER_SPECS:
        if kind == "conv":
            out[f"{scope}/{name}/kernel"] = (k, k, cin, cout)
        elif kind == "deconv":
            out[f"{scope}/{name}/kernel"] = (k, k, cout, cin)
        else:
            out[f"{scope}/{name}/kernel"] = (cin, cout)
        out[f"{scope}/{name}/bias"] = (cout,)
    return out


def make_weights(seed: int = 0) -> dict[str, np.ndarray]:
    """He-uniform kernels U(+-sqrt(6/fan_in)), biases U(+-0.05), with three output-range tweaks so the
    random net exercises the filter non-degenerately: SCoordNet ch3 -> sigma_z in ~[0.02,0.5];
    OFlowNet `uncertainty` -> sigma_w in ~[0.005,0.05]; OFlowNet `prediction` x20 -> peaked softmax."""
    rng = np.random.default_rng(seed)
    w = {}
    for scope, name, kind, k, cin, cout in LAYER_SPECS:
        fan_in = k * k * cin
        lim = np.sqrt(6.0 / fan_in)
        if kind == "conv":
            shape = (k, k, cin, cout)
        elif kind == "deconv":
            shape = (k, k, cout, cin)
        else:
            shape = (cin, cout)
        kern = rng.uniform(-lim, lim, size=shape).astype(np.float32)
        bias = rng.uniform(-0.05, 0.05, size=(cout,)).astype(np.float32)
        if (scope, name) == ("ScoreNet", "prediction"):
            kern[..., 3] *= 0.5
            bias[3] = -2.3
        if (scope, name) == ("Temporal", "uncertainty"):
            kern *= 0.5
            bias[:] = 0.3
        if (scope, name) == ("Temporal", "prediction"):
            kern *= 20.0
        w[f"{scope}/{name}/kernel"] = kern
        w[f"{scope}/{name}/bias"] = bias
    return w


def save_weights(path: str, weights: dict[str, np.ndarray]) -> None:
    np.savez(path, **{k.replace("/", "__"): v for k, v in weights.items()})


def load_weights(path: str) -> dict[str, np.ndarray]:
    z = np.load(path)
    return {k.replace("__", "/"): z[k] for k in z.files}


def make_sequence(T: int, seed: int = 1, height: int = 480, width: int = 640, max_step: int = 8,
                  noise_frames: bool = False) -> np.ndarray:
    """uint8 [T,H,W,3]: a smooth random texture (low-frequency sinusoids + 8-px blocky noise) seen
    through a window that random-walks by <= max_step px per frame (<= 1 label px; total drift kept
    inside +-24 px of the previous frame so the true flow sits inside the [-4,3] label-px window)."""
    rng = np.random.default_rng(seed)
    margin = 64 + max_step * 4
    Hc, Wc = height + 2 * margin, width + 2 * margin
    yy, xx = np.mgrid[0:Hc, 0:Wc].astype(np.float32)
    canvas = np.zeros((Hc, Wc, 3), np.float32)
    for c in range(3):
        for _ in range(6):
            fx, fy = rng.uniform(0.004, 0.05, 2)
            ph = rng.uniform(0, 2 * np.pi)
            canvas[..., c] += rng.uniform(10, 30) * np.sin(2 * np.pi * (fx * xx + fy * yy) + ph)
    blocks = rng.uniform(-40, 40, size=(Hc // 8 + 1, Wc // 8 + 1, 3)).astype(np.float32)
    canvas += np.kron(blocks, np.ones((8, 8, 1), np.float32))[:Hc, :Wc]
    canvas = np.clip(canvas + 128.0, 0, 255)
    frames = np.empty((T, height, width, 3), np.uint8)
    ox = oy = margin
    for t in range(T):
        if noise_frames and t % 7 == 6:
            frames[t] = rng.integers(0, 256, size=(height, width, 3), dtype=np.uint8)
            continue
        frames[t] = canvas[oy:oy + height, ox:ox + width].astype(np.uint8)
        ox = int(np.clip(ox + rng.integers(-max_step, max_step + 1), margin - 48, margin + 48))
        oy = int(np.clip(oy + rng.integers(-max_step, max_step + 1), margin - 48, margin + 48))
    return frames


def make_transform(seed: int = 2) -> np.ndarray:
    """Random rigid 4x4 (float32), the content of transform.txt BEFORE inversion (KFNet/train.py:49-52)."""
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = q.astype(np.float32)
    T[:3, 3] = rng.uniform(-1, 1, 3).astype(np.float32)
    return T


def make_labels(T: int, seed: int = 3, height: int = 480, width: int = 640) -> np.ndarray:
    """float32 [T,H,W,4] = xyz ~ N(0,1), mask ~ Bernoulli(0.9) (label files: KFNet/train.py:219-222)."""
    rng = np.random.default_rng(seed)
    lab = rng.normal(size=(T, height, width, 4)).astype(np.float32)
    lab[..., 3] = (rng.uniform(size=(T, height, width)) < 0.9).astype(np.float32)
    return lab
