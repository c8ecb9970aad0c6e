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
