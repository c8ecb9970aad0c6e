"""Oracle (fp64) vs golden vectors produced by running the REFERENCE SOURCE on the TF1 shim
(tests/golden/make_golden.py).  Pins the oracle's graph wiring to /root/reference's own code."""
import numpy as np
import torch

from oracle import kfnet_oracle as O
from tests.golden import cases

D = torch.float64
TOL = dict(rtol=1e-9, atol=1e-10)


def t(x):
    return torch.as_tensor(np.asarray(x)).to(D)


def _recursive(images, state, weights):
    c, u = O.scoordnet(images, weights, D)
    f = O.oflow_feat(images, weights, D)
    tc, tu = O.build_oflownet(f[0:1], f[1:2], t(state[0]), t(state[1]), weights, D)
    kc, ku = O.build_kf_coord(tc, tu, c[1:2], u[1:2])
    nis = O.get_nis(c[1:2], u[1:2], tc, tu)
    return dict(feat=f, measure_coord=c[1:2], measure_unc=u[1:2], temp_coord=tc, temp_unc=tu, kf_coord=kc,
                kf_unc=ku, nis=nis, trans_kf_coord=O.apply_transform(kc, t(cases.transform())))


def test_recursive_step_small(golden, weights):
    got = _recursive(cases.small_images(2, seed=5), cases.small_state(6), weights)
    for k, v in got.items():
        np.testing.assert_allclose(v.numpy(), golden["small_rec_" + k], err_msg=k, **TOL)


def test_recursive_step_full_480x640(golden, weights):
    got = _recursive(cases.full_images(), cases.full_state(), weights)
    for k, v in got.items():
        np.testing.assert_allclose(v.numpy(), golden["full_rec_" + k], err_msg=k, **TOL)


def test_batch4_unrolled_filter(golden, weights):
    tc, tu, kc, ku = O.kf_coord_batch(cases.small_images(4, seed=7), weights, D)
    for k, v in dict(temp_coord=tc, temp_unc=tu, kf_coord=kc, kf_unc=ku).items():
        np.testing.assert_allclose(v.numpy(), golden["small_b4_" + k], err_msg=k, **TOL)


def test_coord_volume_and_offsets(golden):
    f1, f2 = cases.volume_case()
    diff, offs = O.coord_volume(t(f1), t(f2), 8)
    np.testing.assert_array_equal(offs.numpy(), golden["volume_offsets"])
    np.testing.assert_allclose(diff.numpy(), golden["volume_diff"], rtol=0, atol=0)


def test_bilinear_sampler_edges(golden):
    imgs, coords = cases.bilinear_case()
    out = O.bilinear_sampler(t(imgs), t(coords))
    np.testing.assert_allclose(out.numpy(), golden["bilinear_out"], rtol=1e-12, atol=1e-13)


def test_eval_sequence_matches_recursive_steps(weights):
    """eval loop (eval.py:77-104) == chained recursive steps; faithful batch-of-2 == de-duplicated."""
    imgs = cases.small_images(4, seed=9)
    Tm = cases.transform()
    a = O.eval_sequence(imgs, weights, Tm, seq_len=500, dtype=D, faithful_redundancy=True)
    b = O.eval_sequence(imgs, weights, Tm, seq_len=500, dtype=D, faithful_redundancy=False)
    for fa, fb in zip(a, b):
        np.testing.assert_allclose(fa.out_map.numpy(), fb.out_map.numpy(), rtol=1e-10, atol=1e-12)
    # frame 0 output is the measurement; re-init period 2 resets at frame 2
    c0, u0 = O.scoordnet(imgs[0:1], weights, D)
    np.testing.assert_allclose(a[0].kf_coord.numpy(), c0.numpy(), rtol=1e-9, atol=1e-12)
    r = O.eval_sequence(imgs, weights, Tm, seq_len=2, dtype=D)
    c2, _ = O.scoordnet(imgs[2:3], weights, D)
    np.testing.assert_allclose(r[2].kf_coord.numpy(), c2.numpy(), rtol=1e-9, atol=1e-12)
    # unrolled 4-frame filter == recursive loop (KFNet.py:102-146 vs :85-100)
    tc, tu, kc, ku = O.kf_coord_batch(imgs, weights, D)
    for i in range(4):
        np.testing.assert_allclose(kc[i].numpy(), b[i].kf_coord[0].numpy(), rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(ku[i].numpy(), b[i].kf_unc[0].numpy(), rtol=1e-9, atol=1e-11)


def test_printed_metrics_match_reference_code():
    """Loss / accuracy (KFNet.CoordLossWithUncertainty as KF_fusion calls it), NIS in-band fraction and median distance
    error (eval.py:10-29): the oracle restatement AND the product's host functions against values produced by the
    reference's own source (tests/golden/make_golden.py::metrics)."""
    import os
    from kfnet_b200 import eval as P
    g = np.load(os.path.join(os.path.dirname(cases.__file__), "kfnet_reference_metrics_on_shim.npz"))
    lc = cases.label_case()
    T = cases.transform()
    loss, acc = O.coord_loss_with_uncertainty(lc["pred"], lc["unc"], lc["gt"], lc["mask"], T)
    assert abs(loss - float(g["loss_masked"])) <= 1e-12 * abs(float(g["loss_masked"])) and acc == float(g["accuracy_masked"])
    loss, acc = O.coord_loss_with_uncertainty(lc["pred"], lc["unc"], lc["gt"], None, T)
    assert abs(loss - float(g["loss_nomask"])) <= 1e-12 * abs(float(g["loss_nomask"])) and acc == float(g["accuracy_nomask"])
    assert O.nis_in_band_fraction(lc["nis"]) == float(g["nis_fraction"])
    world = lc["pred"][0] @ T[:3, :3].T + T[:3, 3]
    gt8, mask8 = lc["gt"][0, ::8, ::8], lc["mask"][0, ::8, ::8]
    med, dmap = O.dist_error(world, gt8, mask8)
    assert abs(med - float(g["dist_median_cm"])) <= 1e-12 and np.allclose(dmap, g["dist_map_cm"], rtol=1e-12, atol=0)
    # product host functions (kfnet_b200/eval.py; fp32 label-grid inputs as eval() feeds them)
    loss_p, acc_p = P.coord_loss_with_uncertainty(lc["pred"][0].astype(np.float32), lc["unc"][0].astype(np.float32),
                                                  gt8.astype(np.float32), mask8.astype(np.float32), T.astype(np.float32))
    assert abs(loss_p - float(g["loss_masked"])) <= 2e-5 * abs(float(g["loss_masked"]))
    assert abs(acc_p - float(g["accuracy_masked"])) <= 1.5 / 96                     # one borderline pixel of 96 at fp32
    assert P.get_NIS_measurement(lc["nis"]) == float(g["nis_fraction"])
    med_p, dmap_p = P.dist_error(world.astype(np.float32), gt8.astype(np.float32), mask8.astype(np.float32))
    assert abs(med_p - float(g["dist_median_cm"])) <= 1e-4 and np.allclose(dmap_p, g["dist_map_cm"], rtol=1e-4, atol=1e-5)

