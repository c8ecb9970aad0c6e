"""CPU checks of the bf16-storage model (oracle/kfnet_quant_oracle.py) that the bf16-mode GPU tests are held to."""
import numpy as np
import torch

from oracle import kfnet_oracle as O
from oracle import kfnet_quant_oracle as Q
from tests.golden import cases
from tests.util import rms_rel_err


def test_rounding_is_bf16_nearest_even():
    x = torch.tensor([1.0, 1.00390625, 1.005859375, 1.01171875, -3.1415926], dtype=torch.float64)
    got = Q.q(x)
    want = x.to(torch.float32).to(torch.bfloat16).to(torch.float64)
    assert torch.equal(got, want) and got.dtype == torch.float64
    assert float(got[1]) == 1.0            # 1 + 2^-8 is a tie: rounds to even (1.0)
    assert float(got[2]) == 1.0078125      # 1 + 1.5 * 2^-8 rounds up to 1 + 2^-7


def test_storage_model_is_close_to_but_distinct_from_the_exact_oracle(weights):
    imgs = cases.small_images(2, seed=5).astype(np.float32)
    c, u = O.scoordnet(imgs, weights, torch.float64)
    cq, uq = Q.scoordnet(imgs, weights)
    e = rms_rel_err(cq, c)
    assert 1e-4 < e < 0.1, e              # bf16 storage costs ~1e-2 here; 0 would mean no rounding was injected
    f, fq = O.oflow_feat(imgs, weights, torch.float64), Q.oflow_feat(imgs, weights)
    ef = rms_rel_err(fq, f)
    assert 1e-4 < ef < 0.1, ef
    assert abs(float(fq.pow(2).sum(-1).sqrt().mean()) - 1.0) < 1e-9      # still l2-normalised in fp64


def test_eval_sequence_under_the_storage_model_restores_the_oracle(weights):
    imgs = cases.small_images(3, seed=9)
    Tm = cases.transform().astype(np.float32)
    before = O.oflownet
    outs = Q.eval_sequence(imgs, weights, Tm, seq_len=500)
    assert O.oflownet is before             # the patch is scoped
    ref = O.eval_sequence(imgs.astype(np.float32), weights, Tm, seq_len=500, dtype=torch.float64)
    assert len(outs) == 3
    # frame 0 is a re-initialisation: output = transformed measurement of the storage-model SCoordNet
    cq, uq = Q.scoordnet(imgs[0:1].astype(np.float32), weights)
    assert rms_rel_err(outs[0].kf_coord, cq) == 0.0
    e = rms_rel_err(outs[0].out_map[..., :3], ref[0].out_map[..., :3])
    assert 1e-4 < e < 0.1, e
