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


def load_label8(path, image_size):
    """The only part of a label eval needs: nearest-neighbour /8 (train.py:259-261) of the raw float32 [H,W,4] file.
    Memory-mapped so that only the 60 touched rows are read instead of the whole 4.9 MB."""
    m = np.memmap(path, dtype="<f4", mode="r", shape=(image_size[0], image_size[1], 4))
    return np.ascontiguousarray(m[::8, ::8])


def _frame_metrics(coords, gt, mask):
    """dist_error (eval.py:17-29) vectorised over a chunk: coords/gt [n,h,w,3], mask [n,h,w,1] -> median cm per frame
    (np.median semantics over the pixels with mask * dist > 0; nan when there is none)."""
    d = (np.sqrt(np.square(coords - gt).sum(-1)) * mask[..., 0]).reshape(coords.shape[0], -1)
    valid = d > 0
    cnt = valid.sum(1)
    dd = np.where(valid, d, np.inf)
    dd.sort(axis=1)
    rows = np.arange(d.shape[0])
    lo, hi = np.maximum(cnt - 1, 0) // 2, cnt // 2
    med = 0.5 * (dd[rows, lo] + dd[rows, np.minimum(hi, d.shape[1] - 1)])
    return np.where(cnt > 0, med, np.nan) * 100.0


def eval(image_list, label_list, transform_file, snapshot, scene="stairs", output_folder="", NIS=False,  # noqa: A001
         mode="fp32", chunk=32, weights=None, engine=None, quiet=False, device=None, stats=None, show=False):
    """KFNet/eval.py:31-164.  Frames are decoded by a thread pool straight into pinned staging memory one chunk ahead of
    the GPU; each chunk is ONE C-ABI call (kfn_eval_stream_host: H2D streamed batch by batch, SCoordNet + tower + OFlowNet
    frame-parallel, the serial warp+Kalman scan, D2H of the output maps and of the tensors the printed line needs), with
    the filter state and the last feature map carried on the device between chunks.  `stats` (optional dict) receives
    frames, seconds and frames_per_s of the loop.  `show` (the reference's --show, eval.py:121,127-160): instead of saving
    coord_<idx>.npy, the twelve panels the reference plots per frame are written as show_<idx>.npz (and as a PNG figure
    when matplotlib is importable) -- this box has no display."""
    import time
    import torch
    from concurrent.futures import ThreadPoolExecutor
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
    H, W = spec.image_size
    n_iter = len(image_paths)
    period = spec.sequence_length
    R, tr = transform[:3, :3], transform[:3, 3]
    dists = [[], [], []]
    save_dir = output_folder if (output_folder and os.path.isdir(output_folder)) else None   # eval.py:122
    # input pipeline (reference: 1-thread TF queue runner, train.py:203-239): ~80-100 PNG decodes/s per core
    # (tools/host_pipeline_bench.py); cv2 / numpy file reads release the GIL
    pool = ThreadPoolExecutor(max_workers=max(2, min(64, (os.cpu_count() or 4))))
    stage = [torch.empty(chunk, H, W, 3, dtype=torch.uint8).pin_memory() for _ in range(2)]
    stage_np = [t.numpy() for t in stage]

    def decode_into(path, buf, k):
        import cv2
        img = cv2.imread(path, cv2.IMREAD_COLOR)
        if img is None:
            raise IOError("cannot read image " + path)
        cv2.cvtColor(img, cv2.COLOR_BGR2RGB, dst=buf[k])          # decode_png(channels=3) order, written into pinned memory

    def plan(ci):
        i0 = ci * chunk
        its = list(range(i0, min(i0 + chunk, n_iter)))
        pairs = [groups[i % len(groups)] for i in its]
        buf = stage_np[ci & 1]
        futs_i = [pool.submit(decode_into, image_paths[p[1]], buf, k) for k, p in enumerate(pairs)]
        need = sorted({f for p in pairs for f in p})
        futs_l = {f: pool.submit(load_label8, label_paths[f], spec.image_size) for f in need}
        return its, pairs, futs_i, futs_l

    # the chunked stream evaluates pair (frame k-1, frame k) of the decoded frames; that is the reference's pair stream
    # (train.py:67-71) whenever every non-re-init iteration's first frame is the previous iteration's second frame
    def consecutive(its, pairs, last_cur):
        prev = last_cur
        for i, p in zip(its, pairs):
            if i % period != 0 and p[0] != prev:
                return False
            prev = p[1]
        return True

    post = ThreadPoolExecutor(max_workers=1)
    post_futs = []

    def postprocess(its, pairs, futs_l, out_map, meas_np, temp, kf, nis):
        lab = {f: fut.result() for f, fut in futs_l.items()}
        gt = np.stack([lab[p[1]][..., :3] for p in pairs])
        mask = np.stack([lab[p[1]][..., 3:4] for p in pairs])
        reinit = np.array([(i % period == 0) for i in its])
        t_meas = meas_np[..., :3] @ R.T + tr
        t_temp = np.where(reinit[:, None, None, None], t_meas, temp[..., :3] @ R.T + tr)
        d_m, d_t, d_kf = _frame_metrics(t_meas, gt, mask), _frame_metrics(t_temp, gt, mask), _frame_metrics(out_map[..., :3], gt, mask)
        for lst, v in zip(dists, (d_m, d_t, d_kf)):
            lst.extend(float(x) for x in v)
        if not quiet:
            for k, i in enumerate(its):
                prev, index = pairs[k]
                # the reference scores the ONE estimated frame against BOTH ground-truth frames of its batch-of-2
                # (KFNet.CoordLossWithUncertainty broadcasts [1,60,80,3] against [2,60,80,3]; KFNet.py:214, train.py:250-255)
                gt2 = np.stack([lab[prev][..., :3], lab[index][..., :3]])
                mask2 = np.stack([lab[prev][..., 3:4], lab[index][..., 3:4]])
                l_m, a_m = coord_loss_with_uncertainty(meas_np[k:k + 1, ..., :3], meas_np[k:k + 1, ..., 3:4], gt2, mask2, transform)
                l_t, a_t = coord_loss_with_uncertainty(temp[k:k + 1, ..., :3], temp[k:k + 1, ..., 3:4], gt2, mask2, transform)
                l_k, a_k = coord_loss_with_uncertainty(kf[k:k + 1, ..., :3], kf[k:k + 1, ..., 3:4], gt2, mask2, transform)
                if reinit[k]:
                    print("Reinitialize at the begin of a new sequence.")
                print("%d, frame %d~%d, l_m = %.3f, l_t = %.3f, l_kf = %.3f, a_m = %.3f, a_t = %.3f, a_kf = %.3f, "
                      "d_m = %.3f, d_t = %.3f, d_kf = %.3f, nis = %.3f"
                      % (i, prev, index, l_m, l_t, l_k, a_m, a_t, a_k, d_m[k], d_t[k], d_kf[k], get_NIS_measurement(nis[k])))
        if show and save_dir:                                                              # eval.py:127-160
            for k, p in enumerate(pairs):
                _show_panels(os.path.join(save_dir, "show_" + str(p[1])), mask[k], gt[k], nis[k], meas_np[k], t_meas[k], temp[k], t_temp[k],
                             kf[k], out_map[k])
        elif save_dir:                                                                     # eval.py:122-126
            for k, p in enumerate(pairs):
                pool.submit(np.save, os.path.join(save_dir, "coord_" + str(p[1]) + ".npy"), out_map[k].astype(np.float32, copy=False))

    n_chunks = (n_iter + chunk - 1) // chunk
    nxt = plan(0) if n_chunks else None
    carry, last_cur, state = None, None, None
    t_start = time.perf_counter()
    t_dec = t_gpu = 0.0                    # where the loop's wall time goes: waiting for the decoder, the C-ABI call (+ drains at the end)
    for ci in range(n_chunks):
        its, pairs, futs_i, futs_l = nxt
        t0 = time.perf_counter()
        for f in futs_i:
            f.result()
        t_dec += time.perf_counter() - t0
        host_imgs = stage[ci & 1][:len(its)]
        nxt = plan(ci + 1) if ci + 1 < n_chunks else None           # decode chunk ci+1 while the GPU runs chunk ci
        t0 = time.perf_counter()
        if consecutive(its, pairs, last_cur):
            out_t, diag, carry = engine.eval_stream_host(host_imgs, its[0], transform, carry, reinit_period=period, nis_gate=NIS,
                                                         want_diag=True, capacity=chunk)
            out_map = out_t.numpy()
            meas_np, temp, kf, nis = (diag[k].numpy() for k in ("measure", "temp", "kf", "nis"))
        else:
            out_map, meas_np, temp, kf, nis, carry = _eval_chunk_general(engine, host_imgs, pairs, its[0], transform, carry, period, NIS)
        t_gpu += time.perf_counter() - t0
        last_cur = pairs[-1][1]
        # metrics, printing and the .npy writes of this chunk run on a worker thread (in chunk order) while the GPU evaluates the next
        # chunk; the staging views are reused by the next call, so the worker gets copies
        post_futs.append(post.submit(postprocess, its, pairs, futs_l, out_map.copy(), meas_np.copy(), temp.copy(), kf.copy(), nis.copy()))
    t0 = time.perf_counter()
    for f in post_futs:
        f.result()
    post.shutdown(wait=True)
    t_met = time.perf_counter() - t0
    t0 = time.perf_counter()
    pool.shutdown(wait=True)
    t_tail = time.perf_counter() - t0
    dt = time.perf_counter() - t_start
    if stats is not None:
        stats.update(frames=n_iter, seconds=dt, frames_per_s=n_iter / dt if dt > 0 else float("nan"),
                     wait_decode_s=t_dec, cabi_call_s=t_gpu, drain_postprocess_s=t_met, drain_saves_s=t_tail)
    if not quiet:
        print("Median dist error: ", *[np.median(d) for d in dists])
        print("Mean dist error: ", *[np.mean(d) for d in dists])
        print("stddev error: ", *[np.std(d) for d in dists])
    return dists


def _show_panels(path, mask, gt, nis, meas, t_meas, temp, t_temp, kf, out_map):
    """The 4 x 3 panel grid of eval.py:131-155 as arrays (+ a PNG when matplotlib is available)."""
    def dmap(c):
        return np.minimum(np.sqrt(np.square(c - gt).sum(-1)) * mask[..., 0] * 100.0, 10)
    panels = {"mask": mask[..., 0], "gt_coords": gt, "nis": np.minimum(nis, 20),
              "measure_uncertainty": np.minimum(meas[..., 3], 255.0), "measure_coord": t_meas, "measure_dist": dmap(t_meas),
              "temp_uncertainty": np.minimum(temp[..., 3], 255.0), "temp_coord": t_temp, "temp_dist": dmap(t_temp),
              "kf_uncertainty": np.minimum(kf[..., 3], 255.0), "kf_coord": out_map[..., :3], "kf_dist": dmap(out_map[..., :3])}
    np.savez_compressed(path + ".npz", **panels)
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except Exception:
        return
    fig, _ = plt.subplots(4, 3)
    for i, (name, a) in enumerate(panels.items()):
        plt.subplot(4, 3, i + 1)
        plt.imshow(np.clip(a, 0, 1) if a.ndim == 3 else a)
        plt.title(name, fontsize=6); plt.axis("off")
    fig.savefig(path + ".png", dpi=120)
    plt.close(fig)


def _eval_chunk_general(engine, host_imgs, pairs, i0, transform, carry, period, NIS):
    """Arbitrary pair lists (image lists that wrap around the scene's ranges): the stage APIs with explicit pair gathers.
    `carry` holds the filter state in its first GH*GW*4 floats exactly like the streaming call."""
    import torch
    dev = torch.device("cuda", engine.device)
    P = engine.GH * engine.GW
    frames = sorted({f for p in pairs for f in p})
    # frames of the chunk that are not the second frame of a pair must be decoded here (rare path)
    cur_pos = {p[1]: k for k, p in enumerate(pairs)}
    imgs = {f: host_imgs[cur_pos[f]] for f in frames if f in cur_pos}
    missing = [f for f in frames if f not in imgs]
    if missing:
        raise ValueError("eval: image list order needs frames %r that are not part of this chunk's stream" % missing[:4])
    order = list(imgs)
    pos = {f: k for k, f in enumerate(order)}
    batch = torch.stack([imgs[f] for f in order]).to(dev)
    measure_all = engine.scoordnet(batch)
    feat_all = engine.oflow_feat(batch)
    prev_idx = torch.tensor([pos[p[0]] for p in pairs], device=dev)
    cur_idx = torch.tensor([pos[p[1]] for p in pairs], device=dev)
    flow, sigw, _, _ = engine.oflownet(feat_all[prev_idx].contiguous(), feat_all[cur_idx].contiguous())
    measure = measure_all[cur_idx].unsqueeze(0).contiguous()
    if carry is None:
        carry = torch.zeros(P * 36, dtype=torch.float32, device=dev)
    state = carry[:P * 4].view(1, engine.GH, engine.GW, 4) if i0 % period != 0 else None
    res = engine.kalman_scan(measure, flow.unsqueeze(0), sigw.unsqueeze(0), transform, state_in=state, first_index=i0,
                             reinit_period=period, nis_gate=NIS, diagnostics=True, want_state=True)
    carry[:P * 4] = res["state"].reshape(-1)
    carry[P * 4:] = feat_all[cur_idx[-1]].reshape(-1)
    out_map, temp, kf, nis = (res[k][0].cpu().numpy() for k in ("out_map", "temp", "kf", "nis"))
    return out_map, measure[0].cpu().numpy(), temp, kf, nis, carry


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--input_folder", default="")
    ap.add_argument("--model_folder", default="")
    ap.add_argument("--output_folder", default="")
    ap.add_argument("--scene", default="")
    ap.add_argument("--NIS", action="store_true")
    ap.add_argument("--show", action="store_true", help="write the reference's 12 diagnostic panels per frame instead of coord_<idx>.npy")
    ap.add_argument("--gpu", type=int, default=0)
    ap.add_argument("--mode", default="fp32", choices=["fp32", "bf16", "bf16x3", "fp16x2"])
    ap.add_argument("--chunk", type=int, default=32)
    ap.add_argument("--quiet", action="store_true", help="no per-frame line / summary (the output .npy files are unchanged)")
    a = ap.parse_args(argv)
    snapshot, _ = get_snapshot(a.model_folder)
    eval(os.path.join(a.input_folder, "image_list.txt"), os.path.join(a.input_folder, "label_list.txt"),
         os.path.join(a.input_folder, "transform.txt"), snapshot, scene=a.scene, output_folder=a.output_folder,
         NIS=a.NIS, mode=a.mode, chunk=a.chunk, device=a.gpu, quiet=a.quiet, show=a.show)


if __name__ == "__main__":
    main()
