"""Size-independent properties of the path at the full 480x640 geometry (BASELINE configs[3] shape, shorter T):
they hold for any weights / inputs, so they check the CUDA path where the CPU oracle is too slow to follow."""
import numpy as np
import pytest
import torch

from tests.golden import cases
from tests.util import gpu_setup

pytestmark = pytest.mark.gpu
T = 12


@pytest.fixture(scope="module")
def staged(weights):
    gpu_setup()
    from kfnet_b200 import synth
    from kfnet_b200.engine import Engine
    eng = Engine(weights, mode="bf16", image_size=(480, 640))
    imgs = torch.from_numpy(synth.make_sequence(T, seed=8, noise_frames=True)).cuda()
    meas = eng.scoordnet(imgs)
    feat = eng.oflow_feat(imgs)
    flow, sigw, _, _ = eng.oflownet(feat[:-1].contiguous(), feat[1:].contiguous())
    z = torch.zeros(1, 60, 80, 2, device="cuda")
    flow_t = torch.cat([z, flow]).unsqueeze(0).contiguous()             # entry t belongs to the pair (t-1, t)
    sigw_t = torch.cat([z[..., 0], sigw]).unsqueeze(0).contiguous()
    return eng, imgs, meas.unsqueeze(0).contiguous(), feat, flow_t, sigw_t


def test_flow_stays_inside_the_search_window(staged):
    _, _, _, feat, flow_t, sigw_t = staged
    assert float(flow_t.min()) >= -4.0 - 1e-5 and float(flow_t.max()) <= 3.0 + 1e-5     # soft-argmax over offsets in [-4,3]^2
    assert bool((sigw_t[:, 1:] > 0).all())
    np.testing.assert_allclose(feat.pow(2).sum(-1).cpu().numpy(), 1.0, atol=2e-3)       # l2-normalised features


def test_posterior_never_exceeds_prior_and_gain_in_unit_interval(staged):
    eng, _, meas, _, flow_t, sigw_t = staged
    Tm = np.eye(4, dtype=np.float32)
    r = eng.kalman_scan(meas, flow_t, sigw_t, Tm, reinit_period=500, diagnostics=True)
    temp, kf = r["temp"][0, 1:], r["kf"][0, 1:]
    assert bool((kf[..., 3] <= temp[..., 3] * (1 + 1e-6)).all())                        # sigma+ <= sigma-
    assert bool((kf[..., 3] <= meas[0, 1:, ..., 3] * (1 + 1e-6)).all())                 # ... and <= sigma_z
    assert bool((r["nis"][0, 1:] >= 0).all())
    # frame 0 (re-init): outputs are the measurement, NIS is zero
    assert torch.equal(r["kf"][0, 0], meas[0, 0]) and float(r["nis"][0, 0].abs().max()) == 0.0
    # identity transform: saved map = (kf xyz, 1/sigma+)
    assert torch.equal(r["out_map"][0, ..., :3], r["kf"][0, ..., :3])
    np.testing.assert_allclose(r["out_map"][0, ..., 3].cpu().numpy(), (1.0 / r["kf"][0, ..., 3]).cpu().numpy(), rtol=1e-6)


def test_chunked_scan_with_carried_state_is_bit_identical(staged):
    eng, _, meas, _, flow_t, sigw_t = staged
    Tm = cases.transform().astype(np.float32)
    whole = eng.kalman_scan(meas, flow_t, sigw_t, Tm, reinit_period=5, diagnostics=True, want_state=True)
    state, outs = None, []
    for a, b in ((0, 3), (3, 4), (4, 9), (9, T)):
        part = eng.kalman_scan(meas[:, a:b].contiguous(), flow_t[:, a:b].contiguous(), sigw_t[:, a:b].contiguous(), Tm,
                               state_in=state, first_index=a, reinit_period=5, want_state=True)
        state = part["state"]
        outs.append(part["out_map"])
    assert torch.equal(torch.cat(outs, 1), whole["out_map"]) and torch.equal(state, whole["state"])


def test_graph_replay_and_batching_are_bit_identical(staged):
    eng, imgs, _, _, _, _ = staged
    Tm = cases.transform().astype(np.float32)
    seq = imgs.unsqueeze(0).contiguous()
    a = eng.eval_sequences(seq, Tm, reinit_period=500, frames_per_batch=5, pairs_per_batch=4).clone()
    out = torch.empty_like(a)
    eng.eval_sequences_graphed(seq, Tm, out, reinit_period=500, frames_per_batch=12, pairs_per_batch=11)
    eng.eval_sequences_graphed(seq, Tm, out, reinit_period=500, frames_per_batch=12, pairs_per_batch=11)   # replay
    assert torch.equal(a, out)
    two = torch.cat([seq, seq]).contiguous()                                              # S = 2 identical sequences
    b = eng.eval_sequences(two, Tm, reinit_period=500, frames_per_batch=7, pairs_per_batch=3)
    assert torch.equal(b[0], a[0]) and torch.equal(b[1], a[0])
    c = eng.eval_sequences(two, Tm, reinit_period=500, frames_per_batch=24, pairs_per_batch=32)   # one SCoordNet batch, tower sub-batches 16 + 8
    assert torch.equal(c[0], a[0]) and torch.equal(c[1], a[0])


def test_host_buffer_entry_point_matches_device_path(staged):
    """kfn_eval_sequences_host (pinned host frames in, pinned host maps out, batch-wise H2D on the library's copy stream)
    == kfn_eval_sequences on device-resident frames, bit for bit; repeated calls reuse the staging buffers safely."""
    eng, imgs, _, _, _, _ = staged
    Tm = cases.transform().astype(np.float32)
    seq = imgs.unsqueeze(0).contiguous()
    ref = eng.eval_sequences(seq, Tm, reinit_period=500, frames_per_batch=5, pairs_per_batch=4).cpu()
    host = seq.cpu().pin_memory()
    for _ in range(3):
        out = eng.eval_sequences_host(host, Tm, reinit_period=500, frames_per_batch=5, pairs_per_batch=4)
        assert torch.equal(out, ref)


def test_streaming_entry_point_in_ragged_chunks_is_bit_identical(staged):
    """kfn_eval_stream_host (what the drop-in eval() calls): the sequence delivered in ragged chunks -- 1 frame, 4 frames, a chunk
    that straddles a re-initialisation, a tail -- with the filter state and the last feature map carried on the device ==
    one kfn_eval_sequences call, bit for bit; the diagnostics it returns are the scan's."""
    eng, imgs, meas, _, flow_t, sigw_t = staged
    Tm = cases.transform().astype(np.float32)
    seq = imgs.unsqueeze(0).contiguous()
    whole = eng.eval_sequences(seq, Tm, reinit_period=5, frames_per_batch=5, pairs_per_batch=4).cpu()[0]
    diag_ref = eng.kalman_scan(meas, flow_t, sigw_t, Tm, reinit_period=5, diagnostics=True)
    host = imgs.cpu().pin_memory()
    carry, outs, temps, nises = None, [], [], []
    for a, b in ((0, 1), (1, 5), (5, 6), (6, 11), (11, T)):
        out, diag, carry = eng.eval_stream_host(host[a:b].contiguous().pin_memory(), a, Tm, carry, reinit_period=5, frames_per_batch=3,
                                                pairs_per_batch=2, want_diag=True, capacity=8)
        outs.append(out.clone()); temps.append(diag["temp"].clone()); nises.append(diag["nis"].clone())
    assert torch.equal(torch.cat(outs), whole)
    assert torch.equal(torch.cat(temps), diag_ref["temp"][0].cpu()) and torch.equal(torch.cat(nises), diag_ref["nis"][0].cpu())
    # without the carry a continuing chunk must be refused, not silently warped from garbage (state_in check)
    from kfnet_b200._lib import KfnError
    with pytest.raises((KfnError, ValueError)):
        eng.kalman_scan(meas[:, 2:4].contiguous(), flow_t[:, 2:4].contiguous(), sigw_t[:, 2:4].contiguous(), Tm, state_in=None, first_index=2,
                        reinit_period=5)


def test_edge_cases_single_frame_every_frame_reinit_and_oversized_batches(staged):
    """T = 1 (only a re-initialisation frame: no pair exists), reinit_period = 1 (every frame is its own measurement), batches larger
    than the job, and S = 3 ragged against the batch size."""
    eng, imgs, meas, _, _, _ = staged
    Tm = cases.transform().astype(np.float32)
    one = eng.eval_sequences(imgs[:1].unsqueeze(0).contiguous(), Tm, reinit_period=500, frames_per_batch=64, pairs_per_batch=64)
    R, tr = torch.from_numpy(Tm[:3, :3]).cuda(), torch.from_numpy(Tm[:3, 3]).cuda()
    want0 = torch.cat([meas[0, :1, ..., :3] @ R.T + tr, 1.0 / meas[0, :1, ..., 3:4]], -1)
    assert torch.allclose(one[0], want0, rtol=1e-5, atol=1e-5)                      # output = transformed measurement, 1 / sigma_z
    every = eng.eval_sequences(imgs[:5].unsqueeze(0).contiguous(), Tm, reinit_period=1, frames_per_batch=64, pairs_per_batch=64)
    want = torch.cat([meas[0, :5, ..., :3] @ R.T + tr, 1.0 / meas[0, :5, ..., 3:4]], -1)
    assert torch.allclose(every[0], want, rtol=1e-5, atol=1e-5)
    three = imgs[:9].reshape(3, 3, 480, 640, 3).contiguous()                        # S = 3 sequences of T = 3, batches of 4 frames / 5 pairs
    out3 = eng.eval_sequences(three, Tm, reinit_period=500, frames_per_batch=4, pairs_per_batch=5)
    for s_ in range(3):
        solo = eng.eval_sequences(three[s_:s_ + 1].contiguous(), Tm, reinit_period=500, frames_per_batch=16, pairs_per_batch=16)
        assert torch.equal(out3[s_], solo[0])
    from kfnet_b200._lib import KfnError
    with pytest.raises((KfnError, AssertionError)):
        eng.eval_sequences(imgs[:0].unsqueeze(0).contiguous(), Tm)                  # empty problem -> KFN_ERR_ARG, not a launch


_SCAN_SNIPPET = r"""
import hashlib, sys
import numpy as np, torch
sys.path.insert(0, %r)
from kfnet_b200.engine import Engine
from kfnet_b200.train import initial_weights
eng = Engine(initial_weights(0), mode="fp32", image_size=(480, 640))
g = torch.Generator().manual_seed(3)
S, T = 2, 11
meas = torch.randn(S, T, 60, 80, 4, generator=g); meas[..., 3] = meas[..., 3].abs() + 0.05
flow = torch.rand(S, T, 60, 80, 2, generator=g) * 9 - 5          # beyond the [-4, 3] window: clipped-neighbour taps at the borders
sig = torch.rand(S, T, 60, 80, generator=g) * 0.2 + 0.01
Tm = np.eye(4, dtype=np.float32); Tm[0, 3] = 0.5; Tm[1, 0] = 0.1
r = eng.kalman_scan(meas.cuda().contiguous(), flow.cuda().contiguous(), sig.cuda().contiguous(), Tm, reinit_period=5, nis_gate=True,
                    diagnostics=True, want_state=True)
h = hashlib.sha256()
for k in ("out_map", "temp", "kf", "nis", "taps", "state"):
    h.update(r[k].cpu().numpy().tobytes())
print("SCANHASH", h.hexdigest())
"""


def test_scan_cluster_sizes_bit_identical():
    """The thread-block-cluster scan (any cluster size) is the single-CTA scan spread over several SMs: same arithmetic in the same
    order, so every output -- maps, priors, posteriors, NIS, taps, carried state -- has identical bits.  The cluster size is read once
    per process (KFN_SCAN_CLUSTER_SIZE / KFN_SCAN_CLUSTER), hence one subprocess per variant."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hashes = {}
    for name, env in (("single CTA", {"KFN_SCAN_CLUSTER": "0"}), ("cluster 5", {"KFN_SCAN_CLUSTER_SIZE": "5"}), ("default cluster", {})):
        e = dict(os.environ)
        e.pop("KFN_SCAN_CLUSTER", None); e.pop("KFN_SCAN_CLUSTER_SIZE", None)
        e.update(env)
        out = subprocess.run([sys.executable, "-c", _SCAN_SNIPPET % root], env=e, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        hashes[name] = [l.split()[1] for l in out.stdout.splitlines() if l.startswith("SCANHASH")][0]
    assert len(set(hashes.values())) == 1, hashes
