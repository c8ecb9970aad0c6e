"""GPU parity of single layers through kfn_conv_layer, both compute modes, vs a plain PyTorch fp32 conv of the
same op (oracle ops on the GPU).  BF16 mode: inputs/weights are bf16-representable and the accumulator is fp32,
so with fp32 output the tcgen05 path must agree with the fp32 reference to accumulation-order noise."""
import zlib

import numpy as np
import pytest
import torch

from oracle import kfnet_oracle as O
from tests.golden import cases
from tests.util import assert_close_rms, feature_on, gpu_setup

pytestmark = pytest.mark.gpu
BF16_ULP = 2.0 ** -8   # a bf16-output layer fed the same bf16 operands: within ONE bf16 ulp of the fp32/fp64 result (rms-relative)

# (key, kind, k, cin, cout, stride, relu, N, IH, IW)  -- shapes chosen to exercise every tile-box / swizzle / tap variant
LAYERS = [
    ("ScoreNet/conv1b", "conv", 3, 64, 64, 1, True, 2, 16, 32),       # KC64 BN64, box 16x8x1... s1 halo via TMA OOB
    ("ScoreNet/conv2a", "conv", 3, 64, 256, 2, True, 2, 16, 32),      # stride 2 parity view, BN256
    ("ScoreNet/conv3a", "conv", 3, 256, 512, 2, True, 3, 8, 16),      # two N tiles, odd batch (OOB in n)
    ("ScoreNet/conv4b", "conv", 3, 1024, 1024, 1, True, 1, 8, 12),    # K = 9216: deep pipeline, 4 N tiles, W=12 -> bw=4
    ("ScoreNet/conv7", "conv", 1, 256, 128, 1, True, 2, 8, 12),
    ("ScoreNet/prediction", "conv", 1, 128, 4, 1, False, 2, 8, 12),   # Cout 4 padded to 16
    ("Temporal/feat2", "conv", 3, 16, 32, 2, True, 2, 16, 32),        # KC16 (32B swizzle)
    ("Temporal/feat3", "conv", 3, 32, 32, 1, True, 2, 8, 16),         # KC32 (64B swizzle)
    ("Temporal/feat6", "conv", 3, 64, 128, 2, True, 2, 16, 24),
    ("Temporal/conv0", "conv", 3, 32, 32, 1, True, 50, 8, 8),         # U-Net 8x8 windows, box 8x8x2
    ("Temporal/conv1a", "conv", 3, 32, 32, 2, True, 50, 8, 8),
    ("Temporal/conv2a", "conv", 3, 32, 64, 2, True, 70, 4, 4),
    ("Temporal/conv2b", "conv", 3, 64, 64, 1, True, 70, 2, 2),
    ("Temporal/conv3a", "conv", 3, 64, 128, 2, True, 300, 2, 2),      # -> 1x1, pruned taps
    ("Temporal/conv3b", "conv", 3, 128, 128, 1, True, 300, 1, 1),     # centre tap only
    ("Temporal/conv6", "conv", 3, 48, 16, 1, True, 50, 8, 8),         # Cin 48 -> 3 chunks of 16
    ("Temporal/prediction", "conv", 3, 16, 1, 1, False, 50, 8, 8),
    ("Temporal/fc1", "dense", 1, 128, 64, 1, True, 300, 1, 1),
    ("Temporal/uncertainty", "dense", 1, 32, 1, 1, False, 300, 1, 1),
]
DECONVS = [("Temporal/upconv2", 128, 64, 300, 1), ("Temporal/upconv1", 64, 32, 70, 2), ("Temporal/upconv0", 32, 16, 50, 4)]


@pytest.fixture(scope="module")
def engines(weights):
    gpu_setup()
    from kfnet_b200.engine import Engine
    return {m: Engine(weights, mode=m, image_size=cases.SMALL_HW) for m in ("fp32", "bf16")}


def _ref_layer(weights, key, kind, x, stride, relu, round_w):
    w = torch.tensor(weights[key + "/kernel"]).cuda()
    b = torch.tensor(weights[key + "/bias"]).cuda()
    if round_w:
        w = w.to(torch.bfloat16).float()
    if kind == "dense":
        w = w.reshape(1, 1, *w.shape)
    return O.conv2d_same(x, w, b, stride, relu)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("spec", LAYERS, ids=[s[0] for s in LAYERS])
def test_conv_layer(engines, weights, mode, spec):
    key, kind, k, cin, cout, stride, relu, N, IH, IW = spec
    g = torch.Generator().manual_seed(zlib.crc32(key.encode()) % 1000)        # deterministic across processes (hash() is salted)
    x = torch.randn(N, IH, IW, cin, generator=g).cuda()
    eng = engines[mode]
    OH, OW = IH // stride, IW // stride
    if mode == "bf16":
        xb = x.to(torch.bfloat16)
        ref = _ref_layer(weights, key, kind, xb.float(), stride, relu, True)
        got = eng.conv_layer(key, xb.contiguous(), (OH, OW), cout, out_fp32=True)
        assert_close_rms(got, ref, 2e-4, key + " bf16 operands / fp32 out")
        got16 = eng.conv_layer(key, xb.contiguous(), (OH, OW), cout, out_fp32=False)
        assert_close_rms(got16.float(), ref, BF16_ULP, key + " bf16 out")
    else:
        ref = _ref_layer(weights, key, kind, x, stride, relu, False)
        got = eng.conv_layer(key, x.contiguous(), (OH, OW), cout)
        assert_close_rms(got, ref, 2e-5, key + " fp32")


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("spec", DECONVS, ids=[s[0] for s in DECONVS])
def test_deconv_phases_and_concat_offset(engines, weights, mode, spec):
    """conv2d_transpose = 4 sub-pixel phases, written into channels [0,Cout) of a wider concat buffer."""
    key, cin, cout, N, n = spec
    g = torch.Generator().manual_seed(5)
    x = torch.randn(N, n, n, cin, generator=g).cuda()
    eng = engines[mode]
    w = torch.tensor(weights[key + "/kernel"]).cuda()
    b = torch.tensor(weights[key + "/bias"]).cuda()
    ctot = cout + 32
    if mode == "bf16":
        x = x.to(torch.bfloat16)
        ref = O.conv2d_transpose_same(x.float(), w.to(torch.bfloat16).float(), b, True)
    else:
        ref = O.conv2d_transpose_same(x, w, b, True)
    out = torch.full((N, 2 * n, 2 * n, ctot), 7.0, dtype=x.dtype, device="cuda")
    for ph in range(4):
        eng.conv_layer(f"{key}#p{ph}", x.contiguous(), (2 * n, 2 * n), ctot, out_coff=0, out=out)
    assert_close_rms(out[..., :cout].float(), ref, BF16_ULP if mode == "bf16" else 2e-5, key)
    assert bool((out[..., cout:] == 7.0).all()), "channels outside [0,Cout) must be untouched"
    # the same conv2d_transpose as ONE GEMM (N = 4*Cout, epilogue scatters column blocks to the phase positions),
    # written at a channel offset of a wider buffer
    out2 = torch.full((N, 2 * n, 2 * n, ctot), 7.0, dtype=x.dtype, device="cuda")
    eng.conv_layer(f"{key}#all", x.contiguous(), (2 * n, 2 * n), ctot, out_coff=16, out=out2)
    assert_close_rms(out2[..., 16:16 + cout].float(), ref, BF16_ULP if mode == "bf16" else 2e-5, key + " merged phases")
    assert bool((out2[..., :16] == 7.0).all()) and bool((out2[..., 16 + cout:] == 7.0).all())


def test_first_layer_im2col_weights(engines, weights):
    """BF16 mode runs the Cin=3 layers as a 1-tap GEMM over the 32-wide im2col row with 0.00625 folded into W."""
    g = torch.Generator().manual_seed(1)
    img = torch.randint(0, 256, (2, 16, 24, 3), generator=g).float().cuda()
    x = (img - 128.0).permute(0, 3, 1, 2)
    cols = torch.nn.functional.unfold(x, 3, padding=1).reshape(2, 3, 9, 16, 24)          # [n, c, tap, h, w]
    cols = cols.permute(0, 3, 4, 2, 1).reshape(2, 16, 24, 27)
    cols = torch.cat([cols, torch.zeros(2, 16, 24, 5, device="cuda")], -1).to(torch.bfloat16).contiguous()
    for key, cout in (("ScoreNet/conv1a", 64), ("Temporal/feat1", 16)):
        w = (torch.tensor(weights[key + "/kernel"]).cuda() * 0.00625).to(torch.bfloat16).float() / 0.00625
        ref = O.conv2d_same(O.preprocess(img), w, torch.tensor(weights[key + "/bias"]).cuda(), 1, True)
        got = engines["bf16"].conv_layer(key, cols, (16, 24), cout, out_fp32=True)
        assert_close_rms(got, ref, 5e-4, key)


@pytest.mark.skipif(not feature_on("KFN_TC_IM2COL_FUSED"), reason="fused u8 first layer not enabled")
@pytest.mark.parametrize("shape", [(2, 16, 24), (1, 480, 640), (3, 40, 136)])
def test_first_layer_fused_from_u8(engines, weights, shape):
    """BF16 mode, default path: the first layers read the u8 image; the 32-wide im2col row (u8 - 128, zero outside the
    image) is built in shared memory by the kernel's producer warps (no im2col tensor in HBM).  Full-size frames take
    the 576+96-thread WIDE layout, small ones the 320+96-thread one; N=3 with 8x8x2-pixel tiles has a ragged batch tile."""
    n, hh, ww = shape
    g = torch.Generator().manual_seed(5 + hh)
    img_u8 = torch.randint(0, 256, (n, hh, ww, 3), generator=g, dtype=torch.uint8).cuda()
    img = img_u8.float()
    cols = torch.nn.functional.unfold((img - 128.0).permute(0, 3, 1, 2), 3, padding=1).reshape(n, 3, 9, hh, ww)
    cols = cols.permute(0, 3, 4, 2, 1).reshape(n, hh, ww, 27)
    cols = torch.cat([cols, torch.zeros(n, hh, ww, 5, device="cuda")], -1).to(torch.bfloat16).contiguous()
    for key, cout in (("ScoreNet/conv1a", 64), ("Temporal/feat1", 16)):
        w = (torch.tensor(weights[key + "/kernel"]).cuda() * 0.00625).to(torch.bfloat16).float() / 0.00625
        ref = O.conv2d_same(O.preprocess(img), w, torch.tensor(weights[key + "/bias"]).cuda(), 1, True)
        got = engines["bf16"].conv_layer(key + "#u8", img_u8, (hh, ww), cout, out_fp32=True)
        assert_close_rms(got, ref, 5e-4, key + "#u8 fp32 out")
        # same operands, same two MMAs per tile as the im2col-tensor + TMA path: bit-identical
        assert torch.equal(got, engines["bf16"].conv_layer(key, cols, (hh, ww), cout, out_fp32=True)), key + " fused vs TMA path"

        got16 = engines["bf16"].conv_layer(key + "#u8", img_u8, (hh, ww), cout)          # bf16 output: coalesced store path
        assert_close_rms(got16.float(), ref, BF16_ULP, key + "#u8 bf16 out")
        # writes stay inside [out_coff, out_coff + cout) of a wider tensor
        wide = torch.full((n, hh, ww, cout + 32), 7.0, dtype=torch.bfloat16, device="cuda")
        engines["bf16"].conv_layer(key + "#u8", img_u8, (hh, ww), cout, out_coff=16, out=wide)
        assert bool((wide[..., :16] == 7.0).all()) and bool((wide[..., 16 + cout:] == 7.0).all())
        assert_close_rms(wide[..., 16:16 + cout].float(), ref, BF16_ULP, key + "#u8 offset")


def test_first_layer_fused_rejected_outside_bf16(engines):
    img = torch.zeros(1, 16, 24, 3, dtype=torch.uint8, device="cuda")
    with pytest.raises(RuntimeError):
        engines["bf16"].conv_layer("ScoreNet/conv1b#u8", img, (16, 24), 64)


@pytest.mark.skipif(not feature_on("KFN_TC_WIDE"), reason="WIDE kernel variant not enabled")
@pytest.mark.parametrize("shape", [(1, 480, 640), (3, 240, 320)])
def test_wide_layout_conv1b_full_frame(engines, weights, shape):
    """conv1b at frame scale takes the WIDE kernel variant (576 threads: four accumulator stages, four epilogue groups
    draining whole tiles round-robin) on top of the halo-tile MMA path; the small LAYERS shapes stay below its size
    threshold, so it is checked here against the same fp32 reference (fp32 and bf16 outputs, N=3: ragged last tiles)."""
    n, hh, ww = shape
    key = "ScoreNet/conv1b"
    g = torch.Generator().manual_seed(11 + n)
    x = torch.randn(n, hh, ww, 64, generator=g).to(torch.bfloat16).cuda().contiguous()
    ref = _ref_layer(weights, key, "conv", x.float(), 1, True, True)
    got = engines["bf16"].conv_layer(key, x, (hh, ww), 64, out_fp32=True)
    assert_close_rms(got, ref, 2e-4, key + " wide fp32 out")
    got16 = engines["bf16"].conv_layer(key, x, (hh, ww), 64)
    assert_close_rms(got16.float(), ref, BF16_ULP, key + " wide bf16 out")


@pytest.mark.skipif(not feature_on("KFN_TC_FUSE_CONV1"), reason="fused conv1a+conv1b kernel is opt-in (KFN_TC_FUSE_CONV1=1) until validated")
@pytest.mark.parametrize("shape", [(2, 16, 24), (1, 480, 640), (3, 48, 136)])
def test_fused_conv1ab(engines, weights, shape):
    """conv1a -> ReLU -> bf16 -> conv1b -> ReLU in one kernel (conv1ab_fused.cu): the 64-channel intermediate never
    leaves the SM.  Same bf16 intermediate and the same MMA order as the two-kernel path, so the result must be
    bit-identical to it, and within bf16 output rounding of the fp32 reference built from the same rounded operands."""
    n, hh, ww = shape
    g = torch.Generator().manual_seed(17 + hh)
    img_u8 = torch.randint(0, 256, (n, hh, ww, 3), generator=g, dtype=torch.uint8).cuda()
    e = engines["bf16"]
    mid = e.conv_layer("ScoreNet/conv1a#u8", img_u8, (hh, ww), 64)                    # bf16 intermediate, two-kernel path
    two = e.conv_layer("ScoreNet/conv1b", mid, (hh, ww), 64)
    fused = e.conv_layer("ScoreNet/conv1ab#u8", img_u8, (hh, ww), 64)
    assert torch.equal(fused, two), "fused conv1a+conv1b differs from the two-kernel path"
    ref = _ref_layer(weights, "ScoreNet/conv1b", "conv", mid.float(), 1, True, True)
    assert_close_rms(fused.float(), ref, 1e-2, "fused conv1ab vs fp32 reference")



@pytest.mark.skipif(not feature_on("KFN_TC_PAIR"), reason="CTA-pair tiles switched off (KFN_TC_PAIR=0)")
@pytest.mark.parametrize("key,cin,cout,stride,N,OH,OW", [
    ("ScoreNet/conv2b", 256, 256, 1, 5, 120, 168),      # 945 M-tiles: odd -> the last pair has a phantom half; ragged batch (bn = 2, N = 5)
    ("ScoreNet/conv3a", 256, 512, 2, 5, 120, 168),      # stride-2 parity view, two N tiles
    ("ScoreNet/conv6", 512, 256, 1, 16, 60, 80),        # the 60x80 grid at batch 16: 600 M-tiles, 4.05 waves of pairs
])
def test_cta_pair_layers(engines, weights, key, cin, cout, stride, N, OH, OW):
    """conv_tc_kernel_pair (tcgen05.mma.cta_group::2, 256-row tiles): layers large enough to take the pair plan (>= 2 tiles per SM) vs a
    PyTorch fp32 conv of the same bf16 operands: one bf16 ulp."""
    g = torch.Generator().manual_seed(zlib.crc32(key.encode()) % 1000 + 7)
    x = (torch.randn(N, OH * stride, OW * stride, cin, generator=g) * 0.5).cuda().to(torch.bfloat16)
    ref = _ref_layer(weights, key, "conv", x.float(), stride, True, True)
    got = engines["bf16"].conv_layer(key, x.contiguous(), (OH, OW), cout)
    assert_close_rms(got.float(), ref, BF16_ULP, key + " CTA-pair tiles, bf16 out")
    out = torch.full((N, OH, OW, cout + 64), 3.0, dtype=torch.bfloat16, device="cuda")          # channel offset inside a wider buffer
    engines["bf16"].conv_layer(key, x.contiguous(), (OH, OW), cout + 64, out_coff=32, out=out)
    assert torch.equal(out[..., 32:32 + cout], got) and bool((out[..., :32] == 3.0).all()) and bool((out[..., 32 + cout:] == 3.0).all())
