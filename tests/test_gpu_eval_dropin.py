"""GPU test of the reference-facing entry point: kfnet_b200.eval.eval(image_list, label_list, transform_file, snapshot)
on real files laid out like the reference expects (PNG frames, raw float32 [480,640,4] labels, transform.txt,
model.ckpt-<step>), checking the saved coord_<idx>.npy maps against the oracle's eval loop."""
import os

import numpy as np
import pytest
import torch

from oracle import kfnet_oracle as O

pytestmark = pytest.mark.gpu


def test_eval_cli_matches_oracle_on_files(tmp_path, weights, capsys):
    import cv2
    from kfnet_b200 import eval as ev
    from kfnet_b200 import synth
    T = 5
    imgs = synth.make_sequence(T, seed=4)
    labels = synth.make_labels(T, seed=5)
    Tfile = synth.make_transform(2)
    d_in, d_out, d_model = tmp_path / "in", tmp_path / "out", tmp_path / "model"
    for d in (d_in, d_out, d_model):
        d.mkdir()
    ipaths, lpaths = [], []
    for t in range(T):
        ip, lp = d_in / f"frame-{t:06d}.color.png", d_in / f"frame-{t:06d}.label.bin"
        assert cv2.imwrite(str(ip), np.ascontiguousarray(imgs[t][..., ::-1]))            # RGB -> BGR for imwrite
        labels[t].astype("<f4").tofile(lp)
        ipaths.append(str(ip)); lpaths.append(str(lp))
    (d_in / "image_list.txt").write_text("\n".join(ipaths) + "\n")
    (d_in / "label_list.txt").write_text("\n".join(lpaths) + "\n")
    np.savetxt(d_in / "transform.txt", Tfile)
    synth.save_weights(str(d_model / "model.ckpt-7.npz"), weights)
    snapshot, step = ev.get_snapshot(str(d_model))
    assert step == 7
    dists = ev.eval(str(d_in / "image_list.txt"), str(d_in / "label_list.txt"), str(d_in / "transform.txt"), snapshot,
                    scene="stairs", output_folder=str(d_out), mode="fp32", chunk=2)
    printed = capsys.readouterr().out
    assert "Reinitialize at the begin of a new sequence." in printed and "0, frame 1~0" in printed and "Median dist error" in printed
    # reference semantics (train.py:67-71, eval.py:94-101): iteration i estimates frame group[1]; i == 0 is the pair [1, 0]
    Tm = ev.get_transform(str(d_in / "transform.txt"))
    ref = O.eval_sequence(imgs.astype(np.float32), weights, Tm, seq_len=500, dtype=torch.float64)
    for t in range(T):
        got = np.load(d_out / f"coord_{t}.npy")
        assert got.shape == (60, 80, 4) and got.dtype == np.float32
        want = ref[t].out_map.numpy()
        s = np.sqrt((want[..., :3] ** 2).mean())
        err = np.max(np.abs(got[..., :3] - want[..., :3]) / np.maximum(np.abs(want[..., :3]), s))
        assert err < 1e-4, (t, err)
        np.testing.assert_allclose(np.log(got[..., 3]), np.log(want[..., 3]), atol=1e-4)
    assert len(dists[0]) == T and all(np.isfinite(dists[2]))
    # --show (eval.py:121,127-160): the twelve panels per frame instead of the .npy
    d_show = tmp_path / "show"; d_show.mkdir()
    ev.eval(str(d_in / "image_list.txt"), str(d_in / "label_list.txt"), str(d_in / "transform.txt"), snapshot, scene="stairs",
            output_folder=str(d_show), mode="fp32", chunk=4, quiet=True, show=True)
    pan = np.load(d_show / "show_3.npz")
    assert len(pan.files) == 12 and pan["kf_coord"].shape == (60, 80, 3) and not (d_show / "coord_3.npy").exists()
    np.testing.assert_allclose(pan["kf_coord"], np.load(d_out / "coord_3.npy")[..., :3], rtol=0, atol=1e-6)
    # the printed per-frame scalars (eval.py:113-118): losses / accuracies score the ONE estimated frame against BOTH labels
    # of the iteration's batch-of-2 (KFNet.py:214 broadcast), distances against the estimated frame's own label
    import re
    rows = [l for l in printed.splitlines() if re.match(r"^\d+, frame ", l)]
    assert len(rows) == T
    for t, line in enumerate(rows):
        vals = dict(re.findall(r"(\w+) = (-?[\d.]+|nan)", line))
        prev = 1 if t == 0 else t - 1
        assert line.startswith(f"{t}, frame {prev}~{t},")
        lab2 = labels[[prev, t]]
        for tag, c, u in (("m", ref[t].measure_coord, ref[t].measure_unc), ("t", ref[t].temp_coord, ref[t].temp_unc),
                          ("kf", ref[t].kf_coord, ref[t].kf_unc)):
            l_ref, a_ref = O.coord_loss_with_uncertainty(c.numpy(), u.numpy(), lab2[..., :3], lab2[..., 3:4], Tm)
            assert abs(float(vals["l_" + tag]) - l_ref) < 2e-3, (t, tag, vals, l_ref)
            assert abs(float(vals["a_" + tag]) - a_ref) < 2e-3, (t, tag, vals, a_ref)
        d_ref, _ = O.dist_error(ref[t].out_map.numpy()[..., :3], labels[t][::8, ::8, :3], labels[t][::8, ::8, 3:4])
        assert abs(float(vals["d_kf"]) - d_ref) < 2e-3 * max(1.0, abs(d_ref)), (t, vals["d_kf"], d_ref)


def test_kf_fusion_16_tuple(tmp_path, weights):
    """Eager KF_fusion (train.py:241-266): same 16 outputs, same shapes, consistent with the recursive API."""
    import cv2
    from kfnet_b200 import KFNetDataSpec, eval as ev, synth
    imgs, labels = synth.make_sequence(3, seed=6), synth.make_labels(3, seed=7)
    ip, lp = [], []
    for t in range(3):
        a, b = tmp_path / f"{t}.png", tmp_path / f"{t}.bin"
        cv2.imwrite(str(a), np.ascontiguousarray(imgs[t][..., ::-1])); labels[t].astype("<f4").tofile(b)
        ip.append(str(a)); lp.append(str(b))
    (tmp_path / "i.txt").write_text("\n".join(ip)); (tmp_path / "l.txt").write_text("\n".join(lp))
    spec = KFNetDataSpec(scene="stairs")
    lc = torch.zeros(1, 60, 80, 3).cuda(); lu = torch.ones(1, 60, 80, 1).cuda()
    out = ev.KF_fusion(str(tmp_path / "i.txt"), str(tmp_path / "l.txt"), np.eye(4, dtype=np.float32), lc, lu, spec, index=1,
                       weights=weights)
    assert len(out) == 16 and out[7] == [0, 1]
    for k, c in zip(range(7), (3, 1, 3, 1, 3, 1, 3)):
        assert tuple(out[k].shape) == (1, 60, 80, c)
    assert tuple(out[14].shape) == (2, 60, 80, 3) and tuple(out[15].shape) == (2, 60, 80, 1)
    assert all(isinstance(v, float) for v in out[8:14])
    c64, u64 = O.scoordnet(imgs[1:2].astype(np.float32), weights, torch.float64)
    assert float((out[0].cpu().double() - c64).abs().max()) < 1e-3
