"""Training-side oracle (SURVEY N3 groundwork): loss forward values pinned to the reference's own methods executed on the
TF1 shim; gradients of the full 4-frame training loss checked against finite differences; Adam / decay formulas."""
import math
import os

import numpy as np
import torch

from oracle import kfnet_oracle as O
from oracle import kfnet_train_oracle as TO
from tests.golden import cases

D = torch.float64


def _golden():
    return np.load(os.path.join(os.path.dirname(cases.__file__), "kfnet_reference_training_on_shim.npz"))


def test_training_losses_match_reference_code(weights):
    g, tc = _golden(), cases.training_case()
    out = TO.training_losses(tc["images"], g["gt_network_frame"], (tc["mask"] == 1.0).astype(np.float64), weights, D)
    for k in ("measure", "temporal", "kf"):
        assert abs(float(out[k + "_loss"]) - float(g[k + "_loss"])) <= 1e-9 * abs(float(g[k + "_loss"])), k
        assert float(out[k + "_accuracy"]) == float(g[k + "_accuracy"]), k
    # the reference moves the labels with ApplyTransform(..., inverse=True): same as the case's own construction
    T = cases.transform()
    gt_net = (tc["gt_world"] - T[:3, 3]) @ np.linalg.inv(T[:3, :3]).T
    np.testing.assert_allclose(g["gt_network_frame"], gt_net, rtol=1e-9, atol=1e-12)


def test_smooth_and_reprojection_losses_match_reference_code(weights):
    g, tc = _golden(), cases.training_case()
    imgs = torch.tensor(tc["images"], dtype=D)
    mc, _ = O.scoordnet(tc["images"], weights, D)
    small = imgs[:, ::8, ::8]
    mask = torch.tensor((tc["mask"] == 1.0).astype(np.float64))
    for got, key in ((TO.smooth_loss(mc, small, mask), "smooth_masked"), (TO.smooth_loss(mc, small), "smooth_nomask"),
                     (TO.reprojection_loss(mc, torch.tensor(tc["pixel_map"]), torch.tensor(tc["poses"]), mask), "reproj_masked"),
                     (TO.reprojection_loss(mc, torch.tensor(tc["pixel_map"]), torch.tensor(tc["poses"])), "reproj_nomask")):
        assert abs(float(got) - float(g[key])) <= 1e-9 * abs(float(g[key])), key


def test_training_gradient_matches_finite_differences(weights):
    """d(total loss)/d(weight) from autograd through SCoordNet, the tower, the difference volume, the U-Net, soft-argmax
    flow, the bilinear warp and the three Kalman updates, against central differences on a handful of scalars."""
    g, tc = _golden(), cases.training_case()
    mask = (tc["mask"] == 1.0).astype(np.float64)
    state = {}
    _, scal, grads = TO.train_step(weights, tc["images"], g["gt_network_frame"], mask, state, weight_decay=1e-4)
    assert state["t"] == 1 and math.isfinite(scal["loss"])

    def total(w):
        out = TO.training_losses(tc["images"], g["gt_network_frame"], mask, w, D)
        return float(out["loss"] + TO.l2_regularization({k: torch.tensor(np.asarray(v), dtype=D) for k, v in w.items()}, 1e-4))

    probes = [("ScoreNet/prediction/bias", (1,)), ("ScoreNet/conv7/kernel", (0, 0, 5, 7)), ("Temporal/feat7/kernel", (1, 1, 3, 2)),
              ("Temporal/prediction/bias", (0,)), ("Temporal/uncertainty/kernel", (4, 0)), ("ScoreNet/conv4b/bias", (17,))]
    for name, idx in probes:
        base = np.asarray(weights[name], np.float64)
        eps = 1e-5
        vals = []
        for sgn in (+1, -1):
            w = dict(weights); a = base.copy(); a[idx] += sgn * eps; w[name] = a
            vals.append(total(w))
        fd = (vals[0] - vals[1]) / (2 * eps)
        an = float(grads[name][idx])
        assert abs(fd - an) <= 2e-4 * max(abs(fd), abs(an)) + 1e-8, (name, idx, fd, an)


def test_adam_and_decay_formulas():
    p = {"w": torch.tensor([1.0, -2.0], dtype=D)}
    gr = {"w": torch.tensor([0.5, -0.25], dtype=D)}
    st = {}
    lr = TO.exponential_decay(1e-4, 40000, 80000, 0.5)
    assert abs(lr - 1e-4 * 0.5 ** 0.5) < 1e-18
    new = TO.adam_step(p, gr, st, 0.1)
    # first step: m = 0.1 g, v = 0.001 g^2, lr_t = lr sqrt(1 - 0.999) / (1 - 0.9)  ->  p - lr_t m / (sqrt(v) + eps)
    lr_t = 0.1 * math.sqrt(1 - 0.999) / (1 - 0.9)
    exp = p["w"] - lr_t * (0.1 * gr["w"]) / ((0.001 * gr["w"] ** 2).sqrt() + 1e-8)
    assert torch.allclose(new["w"], exp, rtol=0, atol=1e-15) and st["t"] == 1
    reg = TO.l2_regularization({"a": torch.tensor([3.0, 4.0], dtype=D)}, 1e-4)
    assert abs(float(reg) - 1e-4 * 25.0 / 2.0) < 1e-18
