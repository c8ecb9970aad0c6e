"""Drop-in for the reference's KFNet/eval.py: `eval(image_list, label_list, transform_file, snapshot)` and the
CLI flags of KFNet/train.py:14-47 that eval reads.  Same inputs (image_list.txt / label_list.txt / transform.txt),
same outputs (`coord_<frame idx>.npy`, float32 [60,80,4] = transformed KF coord + 1/uncertainty, eval.py:121-126)
and the same printed per-frame line (eval.py:113-118).  Underneath, instead of one sess.run per frame, frames are
processed in chunks: SCoordNet + feature tower + OFlowNet batched over the chunk (all frame-parallel),
then one serial warp+Kalman scan kernel with the filter state carried between chunks.

Usage:  python -m kfnet_b200.eval --input_folder D --output_folder O --model_folder M --scene S [--NIS] [--mode bf16]
"""
from __future__ import annotations

import argparse
import glob
import os
import re

import numpy as np

from .spec import KFNetDataSpec

_TEST_RANGES = {  # scene -> [start, end) test ranges (KFNet/train.py:75-141, is_training=False branches)
    "chess": [(0, 1000), (1000, 2000)], "fire": [(0, 1000), (1000, 2000)], "heads": [(0, 1000)],
    "office": [(0, 1000), (1000, 2000), (2000, 3000), (3000, 4000)], "pumpkin": [(0, 1000), (1000, 2000)],
    "redkitchen": [(0, 1000), (1000, 2000), (2000, 3000), (3000, 4000), (4000, 5000)],
    "stairs": [(0, 500), (500, 1000)],
}


def read_lines(filepath):
    """tools/io.py:208-212"""
    with open(filepath) as fin:
        return [line.strip() for line in fin.readlines()]


def get_snapshot(model_dir):
    """tools/io.py:185-196 (without its os.chdir side effect): newest model.ckpt-<step> by trailing number.
    Weight containers of this package are `model.ckpt-<step>.npz` (TF variable names, TF layouts)."""
    files = glob.glob(os.path.join(model_dir, "model.ckpt-*.index")) + glob.glob(os.path.join(model_dir, "model.ckpt-*.npz"))
    if not files:
        return None, 0
    step_of = lambda f: int(re.findall(r"model\.ckpt-(\d+)", os.path.basename(f))[-1])
    best = max(files, key=step_of)
    return os.path.join(model_dir, "model.ckpt-%d" % step_of(best)), step_of(best)


def get_transform(transform_file=None):
    """KFNet/train.py:49-58: inverse of the 4x4 in transform.txt (identity when absent)."""
    if transform_file:
        return np.linalg.inv(np.loadtxt(transform_file, dtype=np.float32)).astype(np.float32)
    return np.eye(4, dtype=np.float32)


def full_groups_in_range(start, end):
    """KFNet/train.py:67-71 (_get_full_groups_in_range): [s+1, s], then [i, i+1]."""
    return [[start + 1, start]] + [[i, i + 1] for i in range(start, end - 1)]


def get_indexes(scene):
    if scene not in _TEST_RANGES:
        raise ValueError("Invalid scene: %r" % (scene,))        # reference prints and exit()s (train.py:142-144)
    groups = []
    for s, e in _TEST_RANGES[scene]:
        groups += full_groups_in_range(s, e)
    return groups


def load_image(path):
    """decode_png(channels=3) -> uint8 [H,W,3] RGB (train.py:215-216)."""
    import cv2
    img = cv2.imread(path, cv2.IMREAD_COLOR)
    if img is None:
        raise IOError("cannot read image " + path)
    return np.ascontiguousarray(img[..., ::-1])


def load_label(path, image_size):
    """decode_raw float32 -> [H,W,4] = xyz + mask (train.py:219-222)."""
    return np.fromfile(path, dtype="<f4").reshape(image_size[0], image_size[1], 4)


def dist_error(coords, gt_coords, mask):
    """eval.py:17-29: median distance (cm) over pixels with mask*dist > 0, and the distance map."""
    dists = np.sqrt(np.square(coords - gt_coords).sum(-1)) * mask[:, :, 0]
    v = dists.reshape(-1)
    v = v[v > 0]
    return (float(np.median(v)) * 100.0 if v.size else float("nan")), dists * 100


def get_NIS_measurement(out_NIS):
    """eval.py:10-15: share of positive NIS values inside (0.0157, 2.706)."""
    a = np.asarray(out_NIS).reshape(-1)
    a = a[a > 0.0]
    return float(((a > 0.0157) & (a < 2.706)).sum()) / float(a.size) if a.size else 0.0


def coord_loss_with_uncertainty(pred, unc, gt, mask, transform, dist_threshold=0.05, min_uncertainty=1e-5):
    """KFNet.CoordLossWithUncertainty (KFNet.py:192-232) on host, label-grid inputs; eval only prints it."""
    pred = pred @ transform[:3, :3].T + transform[:3, 3]
    mask = (mask == 1.0).astype(np.float32)
    d = np.square(pred - gt).sum(-1, keepdims=True)
    u = np.maximum(unc, min_uncertainty)
    loss_map = np.minimum(3.0 * np.log(u) + d / (2.0 * np.square(u)), -2.0)       # sic: tf.minimum (KFNet.py:217)
    valid = mask.sum() + 1.0
    d = d * mask
    loss = float((loss_map * mask).sum() / valid)
    acc = float((valid - np.count_nonzero(np.maximum(d - dist_threshold ** 2, 0))) / valid)
    return loss, acc


def load_snapshot_weights(snapshot):
    """`snapshot` = `<model_folder>/model.ckpt-<step>` as returned by get_snapshot.  Accepts this package's `.npz`
    container (TF variable names, TF layouts) or a TensorFlow checkpoint-V2 (`.index` + `.data-*`, tf_checkpoint.py)."""
    from . import synth
    if snapshot is None:
        raise FileNotFoundError("no model.ckpt-* snapshot found")
    if os.path.exists(snapshot + ".npz"):
        return synth.load_weights(snapshot + ".npz")
    if os.path.exists(snapshot + ".index"):
        from .tf_checkpoint import load_kfnet_weights
        return load_kfnet_weights(snapshot)                 # scopes 'ScoreNet' + 'Temporal' (eval.py:66-68)
    if os.path.exists(snapshot):
        return synth.load_weights(snapshot)
    raise FileNotFoundError("snapshot %s(.npz|.index) not found" % snapshot)


def KF_fusion(image_list, label_list, transform, last_coord, last_uncertainty, spec, index=0, engine=None, weights=None,
              mode="fp32"):
    """Eager counterpart of KFNet/train.py:241-266.  The reference builds this graph once and every `sess.run` pops the
    next pair of the eval stream off a queue; here `index` selects the iteration.  Returns the same 16-tuple:
    (measure_coord, measure_uncertainty, temp_coord, temp_uncertainty, KF_coord, KF_uncertainty, NIS, group_indexes,
     measure_loss, measure_accuracy, temp_loss, temp_accuracy, KF_loss, KF_accuracy, gt_coords_8, masks_8)
    with torch CUDA tensors for the maps ([1,60,80,C]) and python floats for the scalars.  `transform` is the inverted 4x4."""
    import torch
    from .engine import Engine
    from .kfnet import KFNet
    image_paths, label_paths = read_lines(image_list), read_lines(label_list)
    groups = get_indexes(spec.scene)
    group = groups[index % len(groups)]
    if engine is None:
        engine = Engine(weights, mode=mode, image_size=spec.image_size)
    images = torch.from_numpy(np.stack([load_image(image_paths[f]) for f in group]))
    spec2 = KFNetDataSpec(batch_size=2, image_size=spec.image_size, scene=spec.scene)
    net = KFNet(images, spec2, False, False, engine=engine)
    measure_coord, measure_uncertainty = net.GetMeasureCoord2()
    temp_coord, temp_uncertainty, KF_coord, KF_uncertainty = net.GetKFCoordRecursive(last_coord, last_uncertainty)
    NIS = net.GetNIS(measure_coord, measure_uncertainty, temp_coord, temp_uncertainty)
    labs = np.stack([load_label(label_paths[f], spec.image_size)[::8, ::8] for f in group])      # train.py:259-261
    gt8, mask8 = labs[..., :3], labs[..., 3:4]
    T = np.asarray(transform, np.float32)
    scal = []
    for c, u in ((measure_coord, measure_uncertainty), (temp_coord, temp_uncertainty), (KF_coord, KF_uncertainty)):
        # the reference broadcasts the single estimated frame against both ground-truth frames of the batch (KFNet.py:214)
        l, a = coord_loss_with_uncertainty(c.cpu().numpy(), u.cpu().numpy(), gt8, mask8, T)
        scal += [l, a]
    dev = measure_coord.device
    return (measure_coord, measure_uncertainty, temp_coord, temp_uncertainty, KF_coord, KF_uncertainty, NIS, list(group),
            scal[0], scal[1], scal[2], scal[3], scal[4], scal[5], torch.from_numpy(gt8).to(dev), torch.from_numpy(mask8).to(dev))


def eval(image_list, label_list, transform_file, snapshot, scene="stairs", output_folder="", NIS=False,  # noqa: A001
         mode="fp32", chunk=16, weights=None, engine=None, quiet=False, device=None):
    import torch
    from .engine import Engine

    image_paths, label_paths = read_lines(image_list), read_lines(label_list)
    assert len(image_paths) == len(label_paths)                                           # eval.py:37
    spec = KFNetDataSpec()                      # sic (eval.py:39-42): sequence_length stays 500 for every scene
    spec.scene, spec.batch_size, spec.image_num = scene, 2, len(image_paths)
    groups = get_indexes(scene)
    transform = get_transform(transform_file)
    if engine is None:
        engine = Engine(weights if weights is not None else load_snapshot_weights(snapshot), mode=mode,
                        image_size=spec.image_size, device=device)
    dev = torch.device("cuda", engine.device)
    GH, GW = engine.GH, engine.GW
    n_iter = len(image_paths)
    state = None
    dists = [[], [], []]
    # input pipeline (reference: 1-thread TF queue runner, train.py:203-239): PNG decode / label reads of chunk k+1 run
    # on a thread pool while the GPU evaluates chunk k (cv2 and numpy file reads release the GIL)
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=max(2, min(32, (os.cpu_count() or 4))))   # ~80-100 PNG decodes/s per core (tools/host_pipeline_bench.py)

    def plan(i0):
        its = list(range(i0, min(i0 + chunk, n_iter)))
        pairs = [groups[i % len(groups)] for i in its]
        frames = sorted({f for p in pairs for f in p})
        futs_i = [pool.submit(load_image, image_paths[f]) for f in frames]
        futs_l = [pool.submit(load_label, label_paths[p[1]], spec.image_size) for p in pairs]
        return its, pairs, frames, futs_i, futs_l

    starts = list(range(0, n_iter, chunk))
    nxt = plan(starts[0]) if starts else None
    for ci, i0 in enumerate(starts):
        its, pairs, frames, futs_i, futs_l = nxt
        nxt = plan(starts[ci + 1]) if ci + 1 < len(starts) else None
        pos = {f: k for k, f in enumerate(frames)}
        imgs = torch.from_numpy(np.stack([f.result() for f in futs_i])).to(dev)
        measure_all = engine.scoordnet(imgs)                                               # [F,GH,GW,4]
        feat_all = engine.oflow_feat(imgs)
        prev_idx = torch.tensor([pos[p[0]] for p in pairs], device=dev)
        cur_idx = torch.tensor([pos[p[1]] for p in pairs], device=dev)
        flow, sigw, _, _ = engine.oflownet(feat_all[prev_idx].contiguous(), feat_all[cur_idx].contiguous())
        measure = measure_all[cur_idx].unsqueeze(0).contiguous()
        res = engine.kalman_scan(measure, flow.unsqueeze(0), sigw.unsqueeze(0), transform, state_in=state, first_index=i0,
                                 reinit_period=spec.sequence_length, nis_gate=NIS, diagnostics=True, want_state=True)
        state = res["state"]
        out_map, temp, kf, nis = (res[k][0].cpu().numpy() for k in ("out_map", "temp", "kf", "nis"))
        meas_np = measure[0].cpu().numpy()
        for k, i in enumerate(its):
            prev, index = pairs[k]
            lab = futs_l[k].result()[::8, ::8]                                             # nearest-neighbour /8 (train.py:259-261)
            gt, mask = lab[..., :3], lab[..., 3:4]
            reinit = (i % spec.sequence_length == 0)
            R, tr = transform[:3, :3], transform[:3, 3]
            t_meas = meas_np[k, ..., :3] @ R.T + tr
            t_temp = t_meas if reinit else temp[k, ..., :3] @ R.T + tr
            d_m, _ = dist_error(t_meas, gt, mask)
            d_t, _ = dist_error(t_temp, gt, mask)
            d_kf, _ = dist_error(out_map[k, ..., :3], gt, mask)
            for lst, v in zip(dists, (d_m, d_t, d_kf)):
                lst.append(v)
            if not quiet:
                l_m, a_m = coord_loss_with_uncertainty(meas_np[k, ..., :3], meas_np[k, ..., 3:4], gt, mask, transform)
                l_t, a_t = coord_loss_with_uncertainty(temp[k, ..., :3], temp[k, ..., 3:4], gt, mask, transform)
                l_k, a_k = coord_loss_with_uncertainty(kf[k, ..., :3], kf[k, ..., 3:4], gt, mask, transform)
                if reinit:
                    print("Reinitialize at the begin of a new sequence.")
                print("%d, frame %d~%d, l_m = %.3f, l_t = %.3f, l_kf = %.3f, a_m = %.3f, a_t = %.3f, a_kf = %.3f, "
                      "d_m = %.3f, d_t = %.3f, d_kf = %.3f, nis = %.3f"
                      % (i, prev, index, l_m, l_t, l_k, a_m, a_t, a_k, d_m, d_t, d_kf, get_NIS_measurement(nis[k])))
            if output_folder and os.path.isdir(output_folder):                             # eval.py:122-126
                np.save(os.path.join(output_folder, "coord_" + str(index) + ".npy"), out_map[k].astype(np.float32))
    pool.shutdown(wait=False)
    if not quiet:
        print("Median dist error: ", *[np.median(d) for d in dists])
        print("Mean dist error: ", *[np.mean(d) for d in dists])
        print("stddev error: ", *[np.std(d) for d in dists])
    return dists


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--input_folder", default="")
    ap.add_argument("--model_folder", default="")
    ap.add_argument("--output_folder", default="")
    ap.add_argument("--scene", default="")
    ap.add_argument("--NIS", action="store_true")
    ap.add_argument("--gpu", type=int, default=0)
    ap.add_argument("--mode", default="fp32", choices=["fp32", "bf16", "bf16x3", "fp16x2"])
    ap.add_argument("--chunk", type=int, default=16)
    a = ap.parse_args(argv)
    snapshot, _ = get_snapshot(a.model_folder)
    eval(os.path.join(a.input_folder, "image_list.txt"), os.path.join(a.input_folder, "label_list.txt"),
         os.path.join(a.input_folder, "transform.txt"), snapshot, scene=a.scene, output_folder=a.output_folder,
         NIS=a.NIS, mode=a.mode, chunk=a.chunk, device=a.gpu)


if __name__ == "__main__":
    main()
