"""GPU parity: memory-bound kernels (difference volume, softmax-flow, warp + Kalman scan) through the C ABI
vs the CPU oracle on the same seeded inputs.  Integer outputs bit-exact; fp32 within the stated tolerance."""
import numpy as np
import pytest
import torch

from oracle import kfnet_oracle as O
from tests.golden import cases
from tests.util import assert_close_rms, gpu_setup

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(weights):
    gpu_setup()
    from kfnet_b200.engine import Engine
    return Engine(weights, mode="fp32", image_size=cases.SMALL_HW)


def _feat(n, gh, gw, seed):
    g = torch.Generator().manual_seed(seed)
    f = torch.randn(n, gh, gw, 32, generator=g)
    return O.l2_normalize(f)


@pytest.mark.parametrize("gh,gw,npairs", [(8, 12, 1), (5, 7, 2), (60, 80, 2), (1, 1, 1)])
def test_diff_volume_fp32_is_bit_exact(eng, gh, gw, npairs):
    f = _feat(npairs + 1, gh, gw, 3)
    got = eng.diff_volume(f[:-1].cuda().contiguous(), f[1:].cuda().contiguous()).cpu()
    for n in range(npairs):
        ref, _ = O.coord_volume(f[n:n + 1], f[n + 1:n + 2], 8)
        ref = ref.reshape(gh * gw, 8, 8, 32)
        assert torch.equal(got[n * gh * gw:(n + 1) * gh * gw], ref)      # one fp32 subtract: exact


def test_diff_volume_bf16_is_rounded_fp32(eng):
    f = _feat(2, 60, 80, 4)
    a = eng.diff_volume(f[:1].cuda(), f[1:].cuda()).cpu()
    b = eng.diff_volume(f[:1].cuda(), f[1:].cuda(), bf16=True).cpu()
    assert torch.equal(b, a.to(torch.bfloat16))


def test_kalman_scan_matches_oracle_sequence(eng):
    """S=2 sequences, T=5 frames, re-init period 3, diagnostics on: fp32 maps + int32 taps vs the oracle loop."""
    S, T, GH, GW = 2, 5, 60, 80
    g = torch.Generator().manual_seed(7)
    meas = torch.randn(S, T, GH, GW, 4, generator=g)
    meas[..., 3] = torch.rand(S, T, GH, GW, generator=g) * 0.3 + 0.02
    flow = (torch.rand(S, T, GH, GW, 2, generator=g) * 7 - 4)
    flow[0, 1, :4, :4] = torch.tensor([1.0, -2.0])          # exact-integer flows (edge quirk x = W-1 etc.)
    flow[0, 2, :, -1, 0] = 0.0
    sw = torch.rand(S, T, GH, GW, generator=g) * 0.04 + 0.005
    Tm = cases.transform().astype(np.float32)
    res = eng.kalman_scan(meas.cuda(), flow.cuda(), sw.cuda(), Tm, reinit_period=3, diagnostics=True, want_state=True)
    res = {k: (v.cpu() if v is not None else None) for k, v in res.items()}
    D = torch.float32
    grid = O.get_pixel_map(1, GH, GW, D)
    worst = 0.0
    for s in range(S):
        lc = lu = None
        for t in range(T):
            z, zu = meas[s, t:t + 1, ..., :3], meas[s, t:t + 1, ..., 3:4]
            if t % 3 == 0:
                tc, tu, kc, ku, nis = z, zu, z, zu, torch.zeros_like(z)
                taps = None
            else:
                tc, tu, taps = O.process_predict(grid + flow[s, t:t + 1], sw[s, t:t + 1].unsqueeze(-1), lc, lu, True)
                kc, ku = O.build_kf_coord(tc, tu, z, zu)
                nis = O.get_nis(z, zu, tc, tu)
            out = torch.cat([O.apply_transform(kc, torch.tensor(Tm)), 1.0 / ku], -1)
            worst = max(worst, assert_close_rms(res["temp"][s, t], torch.cat([tc, tu], -1)[0], 2e-6, "temp"))
            worst = max(worst, assert_close_rms(res["kf"][s, t], torch.cat([kc, ku], -1)[0], 2e-6, "kf"))
            worst = max(worst, assert_close_rms(res["out_map"][s, t], out[0], 5e-6, "out_map"))
            assert_close_rms(res["nis"][s, t], nis[0], 2e-5, "nis")
            if taps is not None:
                assert torch.equal(res["taps"][s, t], taps[0]), "bilinear tap indices must be bit-exact"
            lc, lu = kc, ku
        assert_close_rms(res["state"][s], torch.cat([lc, lu], -1)[0], 2e-6, "state_out")
    print("kalman scan worst rms-rel err", worst)


def test_kalman_nis_gate_and_single_step(eng):
    GH, GW = 8, 12
    g = torch.Generator().manual_seed(9)
    last = torch.ones(1, GH, GW, 4) * torch.tensor([0.5, -1.0, 2.0, 0.05])      # constant state: the warp returns it (off-border)
    noise = torch.randn(1, 1, GH, GW, 3, generator=g)
    noise[..., : GW // 2, :] *= 0.01                                             # left half: consistent measurement
    noise[..., GW // 2:, :] *= 3.0                                               # right half: outliers -> gated
    meas = torch.cat([last[None, ..., :3] + noise, torch.full((1, 1, GH, GW, 1), 0.05)], -1)
    flow = torch.rand(1, 1, GH, GW, 2, generator=g) * 2 - 1
    sw = torch.full((1, 1, GH, GW), 0.01)
    Tm = cases.transform().astype(np.float32)
    a = eng.kalman_scan(meas.cuda(), flow.cuda(), sw.cuda(), Tm, state_in=last.cuda(), first_index=1, reinit_period=0,
                        nis_gate=False, diagnostics=True)
    b = eng.kalman_scan(meas.cuda(), flow.cuda(), sw.cuda(), Tm, state_in=last.cuda(), first_index=1, reinit_period=0,
                        nis_gate=True, diagnostics=True)
    gate = (a["nis"].sum(-1) > 7.815)[0, 0].cpu()
    assert gate.any() and (~gate).any()
    tz = O.apply_transform(meas[0, :, ..., :3], torch.tensor(Tm))[0]
    ob, oa = b["out_map"][0, 0].cpu(), a["out_map"][0, 0].cpu()
    assert_close_rms(ob[..., :3][gate], tz[gate], 5e-6, "gated pixels take the transformed measurement")
    assert torch.equal(ob[..., :3][~gate], oa[..., :3][~gate]) and torch.equal(ob[..., 3], oa[..., 3])
    assert torch.equal(a["kf"], b["kf"])                     # state unaffected by the gate (eval.py:103-104)


def test_softmax_flow_and_sigma_via_oflownet(eng, weights):
    """fp32-mode OFlowNet on one small pair: logits / flow / sigma_w / argmax vs the fp32 and fp64 oracle."""
    imgs = cases.small_images(2, seed=5)
    f32 = O.oflow_feat(imgs, weights, torch.float32)
    flow, sig, logits, amax = eng.oflownet(f32[:1].cuda().contiguous(), f32[1:].cuda().contiguous(), True, True)
    c0 = torch.zeros(1, 8, 12, 3); u0 = torch.ones(1, 8, 12, 1)
    _, _, ex = O.build_oflownet(f32[:1], f32[1:], c0, u0, weights, torch.float32, True)
    _, _, ex64 = O.build_oflownet(f32[:1].double(), f32[1:].double(), c0.double(), u0.double(), weights, torch.float64, True)
    e_fp32_oracle = assert_close_rms(ex["logits"], ex64["logits"], 1e-3, "oracle fp32 vs fp64 logits")
    assert_close_rms(logits.cpu(), ex64["logits"], 1e-4, "logits")
    assert_close_rms(flow.cpu(), ex64["flow"], 1e-4, "flow")
    assert_close_rms(torch.log(sig.cpu()).reshape(-1), torch.log(ex64["sigma_w"]).reshape(-1), 1e-4, "log sigma_w")
    top2 = ex64["logits"].topk(2, -1).values
    decided = (top2[:, 0] - top2[:, 1]) > 1e-4
    assert decided.float().mean() > 0.95
    assert torch.equal(amax.cpu().reshape(-1)[decided], ex64["argmax"][decided]), "argmax must be bit-exact off near-ties"
    print("oracle fp32-vs-fp64 logits err", e_fp32_oracle)


def test_kalman_update_and_nis_pointwise(eng):
    """kfn_kalman_update (KFNet.BuildKFCoord / GetNIS without the warp) vs the oracle."""
    g = torch.Generator().manual_seed(3)
    prior = torch.randn(2, 8, 12, 4, generator=g); prior[..., 3] = torch.rand(2, 8, 12, generator=g) * 0.2 + 0.01
    meas = torch.randn(2, 8, 12, 4, generator=g); meas[..., 3] = torch.rand(2, 8, 12, generator=g) * 0.2 + 0.01
    post, nis = eng.kalman_update(prior.cuda().contiguous(), meas.cuda().contiguous(), want_nis=True)
    kc, ku = O.build_kf_coord(prior[..., :3], prior[..., 3:4], meas[..., :3], meas[..., 3:4])
    assert_close_rms(post.cpu(), torch.cat([kc, ku], -1), 2e-6, "posterior")
    assert_close_rms(nis.cpu(), O.get_nis(meas[..., :3], meas[..., 3:4], prior[..., :3], prior[..., 3:4]), 2e-6, "nis")
    assert bool((post[..., 3].cpu() <= prior[..., 3] + 1e-7).all())                 # sigma+ <= sigma-


def test_standalone_sampler_and_transform_vs_reference_golden(golden):
    """kfn_bilinear_sampler / kfn_apply_transform (mirrors of tools/util.py:bilinear_sampler, KFNet/util.py:ApplyTransform)
    against the vectors produced by running the reference's own functions on the TF1 shim."""
    from kfnet_b200 import util as U
    imgs, coords = cases.bilinear_case()
    out = U.bilinear_sampler(torch.tensor(imgs, dtype=torch.float32).cuda(), torch.tensor(coords, dtype=torch.float32).cuda())
    assert_close_rms(out.cpu(), golden["bilinear_out"], 5e-6, "bilinear_sampler incl. edge behaviour")
    kc = torch.tensor(golden["full_rec_kf_coord"], dtype=torch.float32).cuda()
    tr = U.ApplyTransform(kc, cases.transform())
    assert_close_rms(tr.cpu(), golden["full_rec_trans_kf_coord"], 2e-6, "ApplyTransform")
    pm = U.GetPixelMap(2, 3, 4)
    assert pm.shape == (2, 3, 4, 2) and pm[0, 2, 3].tolist() == [3.0, 2.0]
    assert U.HomoCoord(kc).shape[-1] == 4
