"""GPU parity of the network stages and the whole eval path vs the CPU oracle and the committed golden
vectors (reference source on the TF1 shim).  fp32 / bf16x3 / fp16x2 modes: 1e-4 on coordinate maps and log-sigma
(north_star tolerance).  bf16 mode: no free-hand bound.  (1) Per layer (test_gpu_conv_layers.py): fed the same bf16 inputs,
a kernel's bf16 output is within ONE bf16 ulp (2^-8) of the fp64 result.  (2) End to end, rounding decisions cascade
chaotically through the stack, so two correct bf16 implementations do not agree bit for bit; the yardstick is the
bf16-STORAGE MODEL of the fp64 oracle (oracle/kfnet_quant_oracle.py: exact arithmetic, weights / activations rounded to
bf16 at the same places): the kernels' error against the EXACT oracle must not exceed the model's own error against the exact
oracle by more than a factor (2 on a max-norm, 1.5 on a 99th percentile) -- i.e. the kernels lose nothing beyond what bf16
storage itself costs.  Both numbers are printed."""
import numpy as np
import pytest
import torch

from oracle import kfnet_oracle as O
from oracle import kfnet_quant_oracle as Q
from tests.golden import cases
from tests.util import assert_close_rms, gpu_setup, rel_err_quantile, rms_rel_err

pytestmark = pytest.mark.gpu
BF16_MAX_FACTOR, BF16_P99_FACTOR, BF16_FLOOR = 2.0, 1.5, 1e-3   # kernel error <= factor * storage-model error + floor (both vs the exact oracle)


def bf16_bound(kernel_err, model_err, what, factor=BF16_MAX_FACTOR):
    assert kernel_err <= factor * model_err + BF16_FLOOR, f"bf16 {what}: kernel error {kernel_err:.3e} vs bf16-storage-model error {model_err:.3e}"
    return f"{what}: kernel {kernel_err:.2e} / model {model_err:.2e}"


@pytest.fixture(scope="module")
def small(weights):
    gpu_setup()
    from kfnet_b200.engine import Engine
    return {m: Engine(weights, mode=m, image_size=cases.SMALL_HW) for m in ("fp32", "bf16", "bf16x3", "fp16x2")}


@pytest.fixture(scope="module")
def full(weights):
    gpu_setup()
    from kfnet_b200.engine import Engine
    return {m: Engine(weights, mode=m, image_size=(480, 640)) for m in ("fp32", "bf16", "bf16x3", "fp16x2")}


def _u8(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(torch.uint8).cuda().contiguous()


def test_scoordnet_and_tower_small_vs_golden(small, golden, weights):
    imgs = cases.small_images(2, seed=5)
    m = small["fp32"].scoordnet(_u8(imgs)).cpu()
    f = small["fp32"].oflow_feat(_u8(imgs)).cpu()
    assert_close_rms(m[1:2, ..., :3], golden["small_rec_measure_coord"], 1e-4, "measure coord")
    assert_close_rms(torch.log(m[1:2, ..., 3:4]), np.log(golden["small_rec_measure_unc"]), 1e-4, "log sigma_z")
    assert_close_rms(f, golden["small_rec_feat"], 1e-4, "tower features")
    mb = small["bf16"].scoordnet(_u8(imgs)).cpu()
    fb = small["bf16"].oflow_feat(_u8(imgs)).cpu()
    imf = imgs.astype(np.float32)
    c64, u64 = O.scoordnet(imf, weights, torch.float64)
    f64 = O.oflow_feat(imf, weights, torch.float64)
    cq, uq = Q.scoordnet(imf, weights)
    fq = Q.oflow_feat(imf, weights)
    print("bf16 small:", bf16_bound(rms_rel_err(mb[..., :3], c64), rms_rel_err(cq, c64), "coord"),
          bf16_bound(rms_rel_err(torch.log(mb[..., 3:4]), torch.log(u64)), rms_rel_err(torch.log(uq), torch.log(u64)), "log sigma"),
          bf16_bound(rms_rel_err(fb, f64), rms_rel_err(fq, f64), "feat"))


def test_recursive_step_full_480x640_vs_golden(full, golden, weights):
    """One GetKFCoordRecursive step at full size: every fetched tensor of eval.py:78-83 vs the golden fixture."""
    from kfnet_b200 import KFNet, KFNetDataSpec
    imgs = cases.full_images()
    lc, lu = (torch.tensor(s, dtype=torch.float32).cuda() for s in cases.full_state())
    # bf16 mode: the same step under the bf16-storage model (fp64 arithmetic, bf16 storage) gives the yardstick errors
    imf = imgs.astype(np.float32)
    q_c, q_u = Q.scoordnet(imf[1:2], weights)
    q_f = Q.oflow_feat(imf, weights)
    with Q._patched_oflownet():
        q_tc, q_tu = O.build_oflownet(q_f[0:1], q_f[1:2], lc.cpu().double(), lu.cpu().double(), weights, torch.float64)
    q_kc, q_ku = O.build_kf_coord(q_tc, q_tu, q_c, q_u)
    for mode, tol in (("fp32", 1e-4), ("bf16x3", 1e-4), ("fp16x2", 1e-4), ("bf16", None)):
        net = KFNet(_u8(imgs), KFNetDataSpec(batch_size=2), engine=full[mode])
        mc, mu = net.GetMeasureCoord2()
        tc, tu, kc, ku = net.GetKFCoordRecursive(lc, lu)
        nis = net.GetNIS(mc, mu, tc, tu)
        got = dict(feat=net.temp_feat_maps, measure_coord=mc, log_measure_unc=torch.log(mu), temp_coord=tc, log_temp_unc=torch.log(tu),
                   kf_coord=kc, log_kf_unc=torch.log(ku))
        want = dict(feat=golden["full_rec_feat"], measure_coord=golden["full_rec_measure_coord"], log_measure_unc=np.log(golden["full_rec_measure_unc"]),
                    temp_coord=golden["full_rec_temp_coord"], log_temp_unc=np.log(golden["full_rec_temp_unc"]),
                    kf_coord=golden["full_rec_kf_coord"], log_kf_unc=np.log(golden["full_rec_kf_unc"]))
        errs = {k: rms_rel_err(got[k], want[k]) for k in got}
        print(mode, {k: f"{v:.2e}" for k, v in errs.items()})
        if mode == "bf16":
            model = dict(feat=q_f, measure_coord=q_c, log_measure_unc=torch.log(q_u), temp_coord=q_tc, log_temp_unc=torch.log(q_tu),
                         kf_coord=q_kc, log_kf_unc=torch.log(q_ku))
            for k in ("feat", "measure_coord", "log_measure_unc"):
                print(bf16_bound(errs[k], rms_rel_err(model[k], want[k]), k))
            # warped maps: a sample point crossing the sampler's x = W-1 / y = H-1 edge flips single pixels (in the model too), so the
            # comparison is on the 99th percentile (test_bf16_flow_indices_vs_fp32_mode covers the integer decisions)
            for k in ("temp_coord", "kf_coord", "log_temp_unc", "log_kf_unc"):
                print(bf16_bound(rel_err_quantile(got[k], want[k], 0.99), rel_err_quantile(model[k], want[k], 0.99), k + " p99", BF16_P99_FACTOR))
            continue
        for k, v in errs.items():
            assert v <= tol, f"{mode} {k}: {v:.3e} > {tol}"
        assert_close_rms(nis, golden["full_rec_nis"], 2e-3, "nis")


def test_batch4_unrolled_equals_recursive_and_golden(small, golden):
    """GetKFCoordBatch (KFNet.py:102-146) == scan over the same 4 frames == golden."""
    from kfnet_b200 import KFNet, KFNetDataSpec
    imgs = cases.small_images(4, seed=7)
    net = KFNet(_u8(imgs), KFNetDataSpec(batch_size=4, image_size=cases.SMALL_HW), engine=small["fp32"])
    tc, tu, kc, ku = net.GetKFCoordBatch()
    assert_close_rms(tc, golden["small_b4_temp_coord"], 1e-4, "temp_coord")
    assert_close_rms(kc, golden["small_b4_kf_coord"], 1e-4, "kf_coord")
    assert_close_rms(torch.log(ku), np.log(golden["small_b4_kf_unc"]), 1e-4, "log kf_unc")
    assert_close_rms(torch.log(tu), np.log(golden["small_b4_temp_unc"]), 1e-4, "log temp_unc")


def test_eval_sequences_small_vs_oracle(small, weights):
    """Whole batched path (S=2 sequences x T=6 frames, ragged batches, re-init every 4) vs the fp64 oracle loop."""
    S, T = 2, 6
    Tm = cases.transform().astype(np.float32)
    seqs = [cases.small_images(T, seed=30 + s) for s in range(S)]
    ref = torch.stack([torch.stack([r.out_map for r in O.eval_sequence(q, weights, Tm, seq_len=4, dtype=torch.float64)])
                       for q in seqs]).float()
    dev = _u8(np.stack(seqs))
    out = small["fp32"].eval_sequences(dev, Tm, reinit_period=4, frames_per_batch=5, pairs_per_batch=3).cpu()
    e = assert_close_rms(out[..., :3], ref[..., :3], 1e-4, "eval out coords")
    e2 = assert_close_rms(torch.log(out[..., 3]), torch.log(ref[..., 3]), 1e-4, "eval out log confidence")
    outb = small["bf16"].eval_sequences(dev, Tm, reinit_period=4, frames_per_batch=4, pairs_per_batch=2).cpu()
    refq = torch.stack([torch.stack([r.out_map for r in Q.eval_sequence(q_, weights, Tm, seq_len=4)]) for q_ in seqs]).float()
    eb = rel_err_quantile(outb[..., :3], ref[..., :3], 0.99)
    print("bf16 eval:", bf16_bound(eb, rel_err_quantile(refq[..., :3], ref[..., :3], 0.99), "out coords p99", BF16_P99_FACTOR),
          f"| max: kernel {rms_rel_err(outb[..., :3], ref[..., :3]):.2e} / model {rms_rel_err(refq[..., :3], ref[..., :3]):.2e}")
    outx = small["bf16x3"].eval_sequences(dev, Tm, reinit_period=4, frames_per_batch=4, pairs_per_batch=5).cpu()
    ex = assert_close_rms(outx[..., :3], ref[..., :3], 1e-4, "bf16x3 (error-compensated tensor-core) eval out coords")
    assert_close_rms(torch.log(outx[..., 3]), torch.log(ref[..., 3]), 1e-4, "bf16x3 eval out log confidence")
    outh = small["fp16x2"].eval_sequences(dev, Tm, reinit_period=4, frames_per_batch=4, pairs_per_batch=5).cpu()
    eh = assert_close_rms(outh[..., :3], ref[..., :3], 1e-4, "fp16x2 (two fp16 limbs, three products) eval out coords")
    assert_close_rms(torch.log(outh[..., 3]), torch.log(ref[..., 3]), 1e-4, "fp16x2 eval out log confidence")
    print(f"eval_sequences small: fp32 err {e:.2e}/{e2:.2e}, bf16x3 err {ex:.2e}, fp16x2 err {eh:.2e}, bf16 err {eb:.2e}")
    # determinism / batch-size independence: different batching gives bit-identical results
    out2 = small["fp32"].eval_sequences(dev, Tm, reinit_period=4, frames_per_batch=12, pairs_per_batch=11).cpu()
    assert torch.equal(out, out2)


def test_scoordnet_batch8_full_vs_oracle(full, weights):
    """BASELINE config 2: SCoordNet, batch 8 synthetic 480x640, vs the fp32/fp64 oracle on two of the frames."""
    from kfnet_b200 import synth
    imgs = synth.make_sequence(8, seed=1)
    dev = _u8(imgs)
    m32 = full["fp32"].scoordnet(dev).cpu()
    mbf = full["bf16"].scoordnet(dev).cpu()
    mx3 = full["bf16x3"].scoordnet(dev).cpu()
    mh2 = full["fp16x2"].scoordnet(dev).cpu()
    for t in (0, 5):
        c64, u64 = O.scoordnet(imgs[t:t + 1].astype(np.float32), weights, torch.float64)
        c32, u32 = O.scoordnet(imgs[t:t + 1].astype(np.float32), weights, torch.float32)
        e_or = rms_rel_err(c32, c64)
        e32 = assert_close_rms(m32[t:t + 1, ..., :3], c64, 1e-4, "fp32 mode coord")
        assert_close_rms(torch.log(m32[t:t + 1, ..., 3:4]), torch.log(u64), 1e-4, "fp32 mode log sigma")
        cq, uq = Q.scoordnet(imgs[t:t + 1].astype(np.float32), weights)
        ebf = rms_rel_err(mbf[t:t + 1, ..., :3], c64)
        print(f"frame {t} bf16:", bf16_bound(ebf, rms_rel_err(cq, c64), "coord"),
              bf16_bound(rms_rel_err(torch.log(mbf[t:t + 1, ..., 3:4]), torch.log(u64)), rms_rel_err(torch.log(uq), torch.log(u64)), "log sigma"))
        ex3 = assert_close_rms(mx3[t:t + 1, ..., :3], c64, 1e-4, "bf16x3 mode coord (tcgen05, error-compensated)")
        assert_close_rms(torch.log(mx3[t:t + 1, ..., 3:4]), torch.log(u64), 1e-4, "bf16x3 mode log sigma")
        eh2 = assert_close_rms(mh2[t:t + 1, ..., :3], c64, 1e-4, "fp16x2 mode coord (tcgen05, two fp16 limbs)")
        assert_close_rms(torch.log(mh2[t:t + 1, ..., 3:4]), torch.log(u64), 1e-4, "fp16x2 mode log sigma")
        print(f"frame {t}: oracle fp32 self-error {e_or:.2e}; fp32 mode {e32:.2e}; bf16x3 mode {ex3:.2e}; fp16x2 mode {eh2:.2e}; bf16 mode {ebf:.2e}")


def test_bf16_flow_indices_vs_fp32_mode(full):
    """argmax / tap indices of the bf16 path vs the fp32 path on a real pair, excluding near-ties."""
    from kfnet_b200 import synth
    dev = _u8(synth.make_sequence(2, seed=1))
    f = full["fp32"].oflow_feat(dev)
    fl32, sg32, lg32, am32 = full["fp32"].oflownet(f[:1].contiguous(), f[1:].contiguous(), True, True)
    fb = full["bf16"].oflow_feat(dev)
    flb, sgb, lgb, amb = full["bf16"].oflownet(fb[:1].contiguous(), fb[1:].contiguous(), True, True)
    top2 = lg32.topk(2, -1).values
    margin = (top2[:, 0] - top2[:, 1])
    decided = margin > 0.25 * lg32.std()
    agree = (am32.reshape(-1)[decided] == amb.reshape(-1)[decided]).float().mean().item()
    print(f"bf16 vs fp32: flow err {rms_rel_err(flb, fl32):.2e}, log-sigma_w err {rms_rel_err(sgb.log(), sg32.log()):.2e}, "
          f"argmax agreement off near-ties {agree:.4f} ({int(decided.sum())} of {decided.numel()} px decided)")
    assert agree > 0.97
