"""Host-side logic of the training drop-in (kfnet_b200/train.py) against the reference's code (KFNet/train.py) -- no GPU."""
import os

import numpy as np

from tests.golden import cases


def test_training_groups_follow_get_indexes():
    """train.py:60-65,75-141: windows [i..i+3] and their time reversal for every i in range(start, end - 3), per range."""
    from kfnet_b200.train import STEPVALUE, training_groups
    g = training_groups("stairs")
    assert len(g) == 4 * (500 - 3) * 2
    assert g[0] == [0, 1, 2, 3] and g[1] == [3, 2, 1, 0]
    assert g[2 * 497 - 1] == [499, 498, 497, 496] and g[2 * 497] == [500, 501, 502, 503]      # no window straddles two recordings
    assert len(training_groups("redkitchen")) == 7 * 997 * 2 and len(training_groups("heads")) == 997 * 2
    assert STEPVALUE == {"chess": 100000, "fire": 30000, "heads": 60000, "office": 100000, "pumpkin": 100000,
                         "redkitchen": 100000, "stairs": 100000}                             # set_stepvalue, train.py:329-345
    import pytest
    with pytest.raises(ValueError):
        training_groups("nowhere")


def test_prepare_window_matches_reference_label_handling():
    """train.py:273-280 executed by the reference's own code on the shim (golden gt_network_frame): nearest /8, mask == 1.0,
    ApplyTransform(gt, get_transform(file), inverse=True)."""
    from kfnet_b200.eval import get_transform
    from kfnet_b200.train import prepare_window
    g = np.load(os.path.join(os.path.dirname(cases.__file__), "kfnet_reference_training_on_shim.npz"))
    tc = cases.training_case()
    T = cases.transform()            # plays `transform` = get_transform(transform_file) of train.py:279 in the golden case
    B, h, w, _ = tc["gt_world"].shape
    full = np.zeros((B, h * 8, w * 8, 4), np.float32)         # full-resolution labels whose [::8, ::8] samples are the case's maps
    mask_vals = np.where(tc["mask"] == 1.0, 1.0, 0.5)          # values other than 1.0 must be masked out
    full[:, ::8, ::8, :3] = tc["gt_world"]
    full[:, ::8, ::8, 3:4] = mask_vals
    gt8, mask8 = prepare_window(full, T)
    assert gt8.dtype == np.float32 and mask8.dtype == np.float32 and gt8.shape == (B, h, w, 3) and mask8.shape == (B, h, w, 1)
    np.testing.assert_allclose(gt8, g["gt_network_frame"], rtol=2e-6, atol=2e-6)
    np.testing.assert_array_equal(mask8, (tc["mask"] == 1.0).astype(np.float32))
    assert get_transform(None).tolist() == np.eye(4).tolist()


def _write_windows(tmp_path, T, hw):
    import cv2
    rng = np.random.default_rng(11)
    H, W = hw
    ip, lp, imgs, labs = [], [], [], []
    for t in range(T):
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        lab = rng.standard_normal((H, W, 4)).astype(np.float32)
        lab[..., 3] = (rng.random((H, W)) > 0.3).astype(np.float32)
        a, b = tmp_path / f"frame-{t:06d}.color.png", tmp_path / f"frame-{t:06d}.label.bin"
        assert cv2.imwrite(str(a), np.ascontiguousarray(img[..., ::-1]))
        lab.astype("<f4").tofile(b)
        ip.append(str(a)); lp.append(str(b)); imgs.append(img); labs.append(lab)
    return ip, lp, np.stack(imgs), np.stack(labs)


def test_window_stream_is_the_reference_queue(tmp_path):
    """train.py:195-239: `string_input_producer(indexes, shuffle=FLAGS.shuffle)` walks the window list epoch after epoch -- in list
    order by default, a fresh permutation per epoch with shuffle -- and every dequeued window is decode_png + decode_raw of its frames;
    the loader threads must deliver windows in queue order with exactly the bytes a direct read gives."""
    from types import SimpleNamespace
    from kfnet_b200.train import prepare_window, window_stream
    T, hw = 6, (16, 24)
    ip, lp, imgs, labs = _write_windows(tmp_path, T, hw)
    spec = SimpleNamespace(image_size=hw)
    Tinv = np.linalg.inv(cases.transform())
    groups = [[0, 1, 2, 3], [3, 2, 1, 0], [1, 2, 3, 4], [4, 3, 2, 1], [2, 3, 4, 5]]
    got = list(window_stream(ip, lp, groups, Tinv, spec, 12))
    assert [g for g, *_ in got] == [groups[i % 5] for i in range(12)]                     # in order, wrapping into the next epoch
    for g, images, gt8, mask8 in got:
        np.testing.assert_array_equal(images, imgs[g])
        e_gt, e_mask = prepare_window(labs[g], Tinv)
        np.testing.assert_array_equal(gt8, e_gt); np.testing.assert_array_equal(mask8, e_mask)
    a = [g for g, *_ in window_stream(ip, lp, groups, Tinv, spec, 15, shuffle=True, seed=3)]
    b = [g for g, *_ in window_stream(ip, lp, groups, Tinv, spec, 15, shuffle=True, seed=3, prefetch=1, workers=1)]
    assert a == b                                                                         # the seed pins the permutations, not the threads
    for e in range(3):                                                                    # every epoch is a permutation of the list
        assert sorted(a[5 * e:5 * e + 5]) == sorted(groups)
    assert a[:5] != groups or a[5:10] != groups
    assert list(window_stream(ip, lp, groups, Tinv, spec, 0)) == [] and list(window_stream(ip, lp, [], Tinv, spec, 3)) == []
    it = window_stream(ip, lp, groups, Tinv, spec, 100)                                   # abandoning the stream stops its loaders
    next(it); it.close()


def test_train_loop_feeds_the_trainer_in_queue_order(tmp_path, monkeypatch):
    """kfnet_b200.train.train() on files with the GPU engine replaced by a recorder: which windows reach `step`, how global_step and
    the snapshot cadence follow train.py:347-420.  (The same loop runs on the real engine in tests/test_gpu_training.py.)"""
    import kfnet_b200.train as tr_mod
    from kfnet_b200 import synth
    T = 5
    ip, lp, imgs, labs = _write_windows(tmp_path, T, (480, 640))
    (tmp_path / "image_list.txt").write_text("\n".join(ip) + "\n"); (tmp_path / "label_list.txt").write_text("\n".join(lp) + "\n")
    np.savetxt(tmp_path / "transform.txt", synth.make_transform(2))
    model = tmp_path / "model"; model.mkdir()
    seen, saved = [], []

    class Recorder:
        def __init__(self, weights, spec=None, fix_flownet=False, device=None, step=0, stepvalue=None, **kw):
            self.global_step, self.stepvalue, self.fix_flownet = int(step), stepvalue, fix_flownet
        def step(self, images, gt8, mask8):
            assert images.dtype == np.uint8 and images.shape == (4, 480, 640, 3) and gt8.shape == (4, 60, 80, 3) and mask8.shape == (4, 60, 80, 1)
            seen.append(images[:, 0, 0, 0].copy())
            self.global_step += 1
            return dict(loss=0.0, lr=1e-4)
        def save(self, folder, step=None):
            saved.append(self.global_step)

    monkeypatch.setattr(tr_mod, "Trainer", Recorder)
    t, hist = tr_mod.train(str(tmp_path / "image_list.txt"), str(tmp_path / "label_list.txt"), str(tmp_path / "transform.txt"), scene="stairs",
                           model_folder=str(model), max_steps=5, weights={}, quiet=True, step=7, snapshot_every=3, fix_flownet=True)
    groups = [[0, 1, 2, 3], [3, 2, 1, 0], [1, 2, 3, 4], [4, 3, 2, 1]]                      # get_indexes windows that fit 5 frames
    assert len(hist) == 5 and t.global_step == 12 and t.stepvalue == 100000 and t.fix_flownet
    for s, g in zip(seen, groups + groups[:1]):
        np.testing.assert_array_equal(s, imgs[g][:, 0, 0, 0])
    assert saved == [9, 12]                                                               # every 3rd global step, and the last step
