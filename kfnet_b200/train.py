"""Drop-in for the training side of the reference's KFNet/train.py: `run` (losses of one 4-frame window, train.py:268-298),
`solver` (L2 regulariser + exponential learning-rate decay + AdamOptimizer, train.py:300-315) and the `train` loop
(train.py:347-420), on the fp32 training path of libkfnet_b200 (kfn_train_begin / kfn_train_step; csrc/train.cu).

Where the reference builds a TF graph once and pulls windows off an input queue, this module evaluates eagerly: one
`Trainer.step(images, gt_coords, masks)` = forward + losses + backward + Adam on one window.

Usage:  python -m kfnet_b200.train --input_folder D --model_folder M --scene S [--max_steps N] [--fix_flownet]
"""
from __future__ import annotations

import argparse
import os

import numpy as np

from .spec import KFNetDataSpec

# flag defaults of KFNet/train.py:14-47
BASE_LR, WEIGHT_DECAY, GAMMA, SNAPSHOT, DISPLAY = 1e-4, 1e-4, 0.5, 5000, 10
STEPVALUE = {"chess": 100000, "fire": 30000, "heads": 60000, "office": 100000, "pumpkin": 100000, "redkitchen": 100000,
             "stairs": 100000}     # set_stepvalue (train.py:329-345); max_steps = 5 * stepvalue (train.py:349)
_TRAIN_RANGES = {  # scene -> [start, end) TRAINING ranges (train.py:75-141, is_training=True branches)
    "chess": [(0, 1000), (1000, 2000), (2000, 3000), (3000, 4000)], "fire": [(0, 1000), (1000, 2000)], "heads": [(0, 1000)],
    "office": [(i, i + 1000) for i in range(0, 6000, 1000)], "pumpkin": [(0, 1000), (1000, 2000), (2000, 3000), (3000, 4000)],
    "redkitchen": [(i, i + 1000) for i in range(0, 7000, 1000)], "stairs": [(0, 500), (500, 1000), (1000, 1500), (1500, 2000)],
}


def training_groups(scene):
    """get_indexes(is_training=True) (train.py:60-65,75-141): for every i in a range, the window [i, i+1, i+2, i+3] AND its
    time reversal [i+3, i+2, i+1, i] (the filter is trained in both directions of time)."""
    if scene not in _TRAIN_RANGES:
        raise ValueError("Invalid scene: %r" % (scene,))        # reference prints and exit()s (train.py:142-144)
    groups = []
    for s, e in _TRAIN_RANGES[scene]:
        for i in range(s, e - 3):
            groups.append([i, i + 1, i + 2, i + 3])
            groups.append([i + 3, i + 2, i + 1, i])
    return groups


def initial_weights(seed=None):
    """`sess.run(init_op)` (train.py:396-399) when no snapshot is given: tf.layers' defaults -- glorot_uniform kernels
    (limit = sqrt(6 / (fan_in + fan_out)), fans taken from the last two dims x the receptive field) and zero biases."""
    from .synth import weight_shapes
    rng = np.random.default_rng(seed)
    w = {}
    for name, shape in weight_shapes().items():
        if name.endswith("/bias"):
            w[name] = np.zeros(shape, np.float32)
            continue
        rec = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
        lim = np.sqrt(6.0 / (shape[-2] * rec + shape[-1] * rec))
        w[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
    return w


def prepare_window(labels_full, transform_inv):
    """train.py:273-280: nearest-neighbour /8 of labels and masks, mask == 1.0, labels moved into the network's frame with
    ApplyTransform(gt, transform, inverse=True) where `transform` = get_transform() = inv(transform.txt) -- i.e. gt is
    multiplied by transform.txt itself.  labels_full: [B,H,W,4] (xyz + mask).  Returns (gt8 [B,h,w,3], mask8 [B,h,w,1]) float32."""
    lab8 = np.asarray(labels_full, np.float32)[:, ::8, ::8]
    T = np.linalg.inv(np.asarray(transform_inv, np.float64))
    gt = lab8[..., :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]
    mask = (lab8[..., 3:4] == 1.0).astype(np.float32)
    return np.ascontiguousarray(gt.astype(np.float32)), np.ascontiguousarray(mask)


class Trainer:
    """`run` + `solver` of train.py on one Engine.  step() returns the scalars the reference prints every DISPLAY steps."""

    def __init__(self, weights, spec=None, train_scoordnet=True, fix_flownet=False, base_lr=BASE_LR, weight_decay=WEIGHT_DECAY,
                 stepvalue=None, gamma=GAMMA, device=None, engine=None, step=0):
        from .engine import Engine
        self.spec = spec or KFNetDataSpec()
        self.engine = engine or Engine(weights, mode="fp32", image_size=self.spec.image_size, device=device)
        self.base_lr, self.weight_decay, self.gamma = base_lr, weight_decay, gamma
        self.stepvalue = stepvalue if stepvalue is not None else STEPVALUE.get(self.spec.scene, 80000)
        self.engine.train_begin(self.spec.batch_size, train_scoordnet, not fix_flownet)   # train.py:283-285
        self.global_step = int(step)
        if self.global_step:
            self.engine.train_set_step(self.global_step)                                  # init_step = global_step.assign(step)

    def step(self, images, gt_coords, masks, apply_update=True):
        """images: uint8 [B,H,W,3]; gt_coords [B,h,w,3] / masks [B,h,w,1] as prepare_window() returns them (numpy or cuda tensors)."""
        import torch
        dev = torch.device("cuda", self.engine.device)
        to = lambda a, dt: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(dev, dt).contiguous()
        out = self.engine.train_step(to(images, torch.uint8), to(gt_coords, torch.float32), to(masks, torch.float32), self.base_lr,
                                     self.weight_decay, self.stepvalue, self.gamma, apply_update)
        if apply_update:
            self.global_step += 1
        return out

    def weights(self):
        from .synth import weight_shapes
        return {name: self.engine.train_get(name) for name in weight_shapes()}

    def save(self, model_folder, step=None):
        """`saver.save(sess, model_folder/model.ckpt, global_step)` (train.py:379-384) as this package's .npz container."""
        from . import synth
        path = os.path.join(model_folder, "model.ckpt-%d.npz" % (self.global_step if step is None else step))
        synth.save_weights(path, self.weights())
        return path


def window_stream(image_paths, label_paths, groups, Tinv, spec, n, shuffle=False, seed=None, prefetch=3, workers=2):
    """The training input queue of train.py:195-239 as a generator of `n` windows `(group, images u8 [B,H,W,3], gt8, mask8)`.

    Order = `tf.train.string_input_producer(indexes, shuffle=FLAGS.shuffle)` (train.py:205): the window list is walked epoch after
    epoch, in list order with `shuffle=False` (the reference's default, train.py:17) and in a fresh random permutation per epoch with
    `shuffle=True` (`seed` pins the permutations; TF draws them from the graph seed).  The reference decodes with ONE queue-runner
    thread and a capacity of 4 batches (train.py:234-239); here PNG decode, the raw-float label reads and `prepare_window` of the next
    `prefetch` windows run on a small thread pool while the GPU executes the current step, and results are delivered in queue order."""
    from collections import deque
    from concurrent.futures import ThreadPoolExecutor
    from .eval import load_image, load_label

    def order():
        rng = np.random.default_rng(seed)
        while True:
            idx = rng.permutation(len(groups)) if shuffle else range(len(groups))
            for gi in idx:
                yield groups[int(gi)]

    def load(g):
        images = np.stack([load_image(image_paths[f]) for f in g])
        labels = np.stack([load_label(label_paths[f], spec.image_size) for f in g])
        gt8, mask8 = prepare_window(labels, Tinv)
        return g, images, gt8, mask8

    if n <= 0 or not groups:
        return
    src = order()
    with ThreadPoolExecutor(max_workers=max(1, int(workers))) as pool:
        pending = deque()
        issued = 0
        while issued < n and len(pending) < max(1, int(prefetch)):
            pending.append(pool.submit(load, next(src))); issued += 1
        while pending:
            item = pending.popleft().result()
            if issued < n:
                pending.append(pool.submit(load, next(src))); issued += 1
            yield item


def train(image_list, label_list, transform_file, snapshot=None, scene="stairs", model_folder="", max_steps=None, fix_flownet=False,
          weights=None, display=DISPLAY, snapshot_every=SNAPSHOT, device=None, quiet=False, step=0, reset_step=-1, scoordnet="",
          oflownet="", shuffle=False, seed=None):
    """The loop of train.py:347-420 over the scene's training windows, fed by `window_stream` (list order like the reference's
    default `--shuffle=False`, or a fresh permutation per epoch).  `step` / `reset_step` set global_step (train.py:361-362);
    `scoordnet` / `oflownet` = model folders whose newest snapshot overrides the 'ScoreNet' / 'Temporal' scope (train.py:402-407);
    with both given the step is reset to 4 x stepvalue like set_stepvalue does (train.py:346-348); the loop ends at
    max_steps = 5 x stepvalue (train.py:349) unless `max_steps` says otherwise."""
    from .eval import get_transform, load_snapshot_weights, read_lines
    image_paths, label_paths = read_lines(image_list), read_lines(label_list)
    assert len(image_paths) == len(label_paths)
    spec = KFNetDataSpec(scene=scene)
    groups = [g for g in training_groups(scene) if max(g) < len(image_paths)]
    if not groups:
        raise ValueError("no training window fits the image list")
    Tinv = get_transform(transform_file)
    from .eval import get_snapshot
    w = dict(weights if weights is not None else (load_snapshot_weights(snapshot) if snapshot else initial_weights()))
    for folder, scope in ((scoordnet, "ScoreNet/"), (oflownet, "Temporal/")):           # RestoreFromScope (train.py:402-407)
        if folder:
            part = load_snapshot_weights(get_snapshot(folder)[0])
            w.update({k: v for k, v in part.items() if k.startswith(scope)})
    stepvalue = STEPVALUE.get(scene, 80000)
    if scoordnet and oflownet and reset_step < 0:
        reset_step = stepvalue * 4
    if reset_step >= 0:
        step = reset_step
    tr = Trainer(w, spec, fix_flownet=fix_flownet, device=device, step=step, stepvalue=stepvalue)
    n = max_steps if max_steps is not None else max(0, stepvalue * 5 - step + 1)           # while step <= FLAGS.max_steps
    hist = []
    for it, (_, images, gt8, mask8) in enumerate(window_stream(image_paths, label_paths, groups, Tinv, spec, n, shuffle=shuffle, seed=seed)):
        out = tr.step(images, gt8, mask8)
        hist.append(out)
        if not quiet and (it % display == 0):
            print("step %d, lr = %.6g, loss = %.4f, m_loss = %.4f, t_loss = %.4f, kf_loss = %.4f, m_acc = %.3f, t_acc = %.3f, kf_acc = %.3f, "
                  "reg = %.4f" % (tr.global_step, out["lr"], out["loss"], out["measure_loss"], out["temporal_loss"], out["kf_loss"],
                                  out["measure_accuracy"], out["temporal_accuracy"], out["kf_accuracy"], out["reg_loss"]))
        if model_folder and os.path.isdir(model_folder) and (tr.global_step % snapshot_every == 0 or it + 1 == n):
            tr.save(model_folder)
    return tr, hist


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--input_folder", default="")
    ap.add_argument("--model_folder", default="")
    ap.add_argument("--scene", default="")
    ap.add_argument("--max_steps", type=int, default=None)
    ap.add_argument("--fix_flownet", action="store_true")
    ap.add_argument("--reset_step", type=int, default=-1)
    ap.add_argument("--scoordnet", default="", help="model folder whose ScoreNet variables are restored over the snapshot")
    ap.add_argument("--oflownet", default="", help="model folder whose Temporal variables are restored over the snapshot")
    ap.add_argument("--gpu", type=int, default=0)
    ap.add_argument("--shuffle", action="store_true", help="FLAGS.shuffle of train.py:17: a fresh permutation of the windows per epoch")
    a = ap.parse_args(argv)
    from .eval import get_snapshot
    snapshot, step = get_snapshot(a.model_folder)
    train(os.path.join(a.input_folder, "image_list.txt"), os.path.join(a.input_folder, "label_list.txt"),
          os.path.join(a.input_folder, "transform.txt"), snapshot, scene=a.scene, model_folder=a.model_folder, max_steps=a.max_steps,
          fix_flownet=a.fix_flownet, device=a.gpu, step=step, reset_step=a.reset_step, scoordnet=a.scoordnet, oflownet=a.oflownet,
          shuffle=a.shuffle)


if __name__ == "__main__":
    main()
