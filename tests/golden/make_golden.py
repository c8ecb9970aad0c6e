"""Generate tests/golden/*.npz by EXECUTING the reference's own model files on oracle/tf1_shim.

Run in the build container only (needs /root/reference):  python -m tests.golden.make_golden
The GPU box never runs this; tests read the committed .npz files.
"""
import builtins
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("KFNET_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
from tests.golden import cases  # noqa: E402

# reference on the shim: shim first, then the reference root (cnn_wrapper, tools) and KFNet/ (util, KFNet)
sys.path.insert(0, os.path.join(REF, "KFNet"))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf1_shim"))
builtins.basestring = str           # cnn_wrapper/network.py:84 is python-2 code

import tensorflow as tf  # noqa: E402  (the shim)
import KFNet as ref_kfnet  # noqa: E402  (/root/reference/KFNet/KFNet.py)
from tools.util import bilinear_sampler as ref_bilinear  # noqa: E402
import util as ref_util  # noqa: E402  (/root/reference/KFNet/util.py)

assert ref_kfnet.__file__.startswith(REF), ref_kfnet.__file__


def npy(t):
    return t.detach().cpu().numpy()


def run_recursive(images, state, out, prefix):
    spec = ref_kfnet.KFNetDataSpec(batch_size=2, image_size=images.shape[1:3])
    net = ref_kfnet.KFNet(tf.constant(images), spec, train_scoordnet=False, train_oflownet=False)
    last_coord, last_unc = (tf.constant(s) for s in state)
    mc, mu = net.GetMeasureCoord2()
    tc, tu, kc, ku = net.GetKFCoordRecursive(last_coord, last_unc)
    nis = net.GetNIS(mc, mu, tc, tu)
    T = tf.constant(cases.transform())
    out.update({prefix + "feat": npy(net.temp_feat_maps), prefix + "measure_coord": npy(mc),
                prefix + "measure_unc": npy(mu), prefix + "temp_coord": npy(tc), prefix + "temp_unc": npy(tu),
                prefix + "kf_coord": npy(kc), prefix + "kf_unc": npy(ku), prefix + "nis": npy(nis),
                prefix + "trans_kf_coord": npy(ref_util.ApplyTransform(kc, T))})
    return net


def main():
    tf.set_dtype(torch.float64)
    tf.set_variables(cases.weights())
    out = {}

    # 1. one recursive filter step at the small size (KFNet.GetKFCoordRecursive, KFNet.py:85-100)
    run_recursive(cases.small_images(2, seed=5), cases.small_state(6), out, "small_rec_")

    # 2. 4-frame unrolled filter (KFNet.GetKFCoordBatch, KFNet.py:102-146)
    imgs4 = cases.small_images(4, seed=7)
    spec = ref_kfnet.KFNetDataSpec(batch_size=4, image_size=imgs4.shape[1:3])
    net = ref_kfnet.KFNet(tf.constant(imgs4), spec, False, False)
    tc, tu, kc, ku = net.GetKFCoordBatch()
    out.update(small_b4_temp_coord=npy(tc), small_b4_temp_unc=npy(tu), small_b4_kf_coord=npy(kc),
               small_b4_kf_unc=npy(ku))

    # 3. difference volume + offsets table (KFNet.BuildCoordVolume, KFNet.py:343-359)
    f1, f2 = cases.volume_case()
    diff, offs = net.BuildCoordVolume(tf.constant(f1), tf.constant(f2), 8)
    out.update(volume_diff=npy(diff), volume_offsets=npy(offs))

    # 4. bilinear sampler incl. edge behaviour (tools/util.py:3-94)
    imgs, coords = cases.bilinear_case()
    out["bilinear_out"] = npy(ref_bilinear(tf.constant(imgs), tf.constant(coords)))

    # 5. full 480x640 recursive step on the synthetic sequence (frames 0,1)
    run_recursive(cases.full_images(), cases.full_state(), out, "full_rec_")

    path = os.path.join(HERE, "kfnet_reference_on_shim.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})
    metrics(net)
    training_losses()


def metrics(net):
    """6. the printed eval metrics: KFNet.CoordLossWithUncertainty exactly as KF_fusion calls it (train.py:254-259:
    full-resolution labels, downsample=True, transform) and eval.py's dist_error / get_NIS_measurement.  eval.py is
    python-2 source (print statements), so the two function definitions are cut out of its text and exec'ed with a
    list-returning `filter` -- the reference's own lines, not a restatement."""
    lc = cases.label_case()
    T = tf.constant(cases.transform())
    out = {}
    for tag, mask in (("masked", tf.constant(lc["mask"])), ("nomask", None)):
        loss, acc = net.CoordLossWithUncertainty(tf.constant(lc["pred"]), tf.constant(lc["unc"]), tf.constant(lc["gt"]),
                                                 mask=mask, transform=T, downsample=True)
        out["loss_" + tag] = np.asarray(float(loss)); out["accuracy_" + tag] = np.asarray(float(acc))
    src = open(os.path.join(REF, "KFNet", "eval.py")).read()
    chunk = src[src.index("def get_NIS_measurement"):src.index("def eval(")]
    ns = {"np": np, "filter": lambda f, xs: list(builtins.filter(f, xs))}
    exec(compile(chunk, os.path.join(REF, "KFNet", "eval.py"), "exec"), ns)
    out["nis_fraction"] = np.asarray(ns["get_NIS_measurement"](lc["nis"]))
    world = npy(ref_util.ApplyTransform(tf.constant(lc["pred"]), T))[0]
    med, dmap = ns["dist_error"](world, lc["gt"][0, ::8, ::8], lc["mask"][0, ::8, ::8])
    out["dist_median_cm"] = np.asarray(med); out["dist_map_cm"] = dmap
    path = os.path.join(HERE, "kfnet_reference_metrics_on_shim.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (v.shape, float(v) if v.ndim == 0 else None) for k, v in out.items()})


def training_losses():
    """7. training-graph losses, executed from the reference's methods the way train.py::run wires them
    (train.py:272-296): labels on the 1/8 grid, gt moved by ApplyTransform(..., inverse=True), mask == 1.0,
    Measure/Temporal/KF CoordLoss with images=resize_images (-> + 50 x SmoothLoss); plus SmoothLoss and
    ReprojectionLoss on their own."""
    tc = cases.training_case()
    imgs = tf.constant(tc["images"])
    spec = ref_kfnet.KFNetDataSpec(batch_size=4, image_size=tc["images"].shape[1:3])
    net = ref_kfnet.KFNet(imgs, spec, train_scoordnet=True, train_oflownet=True)
    h8, w8 = tc["gt_world"].shape[1:3]
    resize_images = tf.image.resize_nearest_neighbor(imgs, [h8, w8])
    masks = tf.cast(tf.equal(tf.constant(tc["mask"]), 1.0), tf.float32)
    T = tf.constant(cases.transform())
    gt = ref_util.ApplyTransform(tf.constant(tc["gt_world"]), T, inverse=True)
    out = {"gt_network_frame": npy(gt)}
    for name, fn in (("measure", net.MeasureCoordLoss), ("temporal", net.TemporalCoordLoss), ("kf", net.KFCoordLoss)):
        loss, acc = fn(gt, masks, images=resize_images)
        out[name + "_loss"] = np.asarray(float(loss)); out[name + "_accuracy"] = np.asarray(float(acc))
    mc, _ = net.GetMeasureCoord()
    out["smooth_masked"] = np.asarray(float(net.SmoothLoss(mc, resize_images, masks)))
    out["smooth_nomask"] = np.asarray(float(net.SmoothLoss(mc, resize_images)))
    out["reproj_masked"] = np.asarray(float(net.ReprojectionLoss(mc, tf.constant(tc["pixel_map"]), tf.constant(tc["poses"]), masks)))
    out["reproj_nomask"] = np.asarray(float(net.ReprojectionLoss(mc, tf.constant(tc["pixel_map"]), tf.constant(tc["poses"]))))
    path = os.path.join(HERE, "kfnet_reference_training_on_shim.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (float(v) if v.ndim == 0 else v.shape) for k, v in out.items()})


if __name__ == "__main__":
    main()
