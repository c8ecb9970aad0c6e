"""Property-based tests (hypothesis) of the oracle's building blocks -- invariants that must hold for any input,
mirrored by the GPU property tests in test_gpu_properties.py."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import kfnet_oracle as O

D = torch.float64
small = settings(max_examples=25, deadline=None)


@small
@given(st.integers(0, 2**31 - 1), st.integers(-2, 2), st.integers(-2, 2))
def test_diff_volume_is_shift_equivariant_in_the_interior(seed, sx, sy):
    """Translating both feature maps translates the volume (away from the zero-filled borders)."""
    rng = np.random.default_rng(seed)
    H, W, C = 14, 15, 3
    f1, f2 = torch.tensor(rng.normal(size=(1, H, W, C))), torch.tensor(rng.normal(size=(1, H, W, C)))
    d, _ = O.coord_volume(f1, f2, 8)
    d_s, _ = O.coord_volume(torch.roll(f1, (sy, sx), (1, 2)), torch.roll(f2, (sy, sx), (1, 2)), 8)
    m = 7                                                     # margin: window reach (4) + shift (2) + 1
    a = torch.roll(d, (sy, sx), (1, 2))[:, m:H - m, m:W - m]
    assert torch.equal(a, d_s[:, m:H - m, m:W - m])


@small
@given(st.integers(0, 2**31 - 1), st.floats(0.1, 30.0))
def test_soft_argmax_flow_stays_in_window_and_tends_to_argmax(seed, sharp):
    rng = np.random.default_rng(seed)
    logits = torch.tensor(rng.normal(size=(6, 64)) * sharp)
    prob = torch.softmax(logits, -1)
    _, offs = O.coord_volume(torch.zeros(1, 1, 1, 1, dtype=D), torch.zeros(1, 1, 1, 1, dtype=D), 8)
    flow, pm = O.flow_from_prob(prob, offs, 1, 2, 3)
    assert float(flow.min()) >= -4 - 1e-12 and float(flow.max()) <= 3 + 1e-12
    if sharp >= 25.0:                                         # peaked softmax -> flow ~ offset of the argmax
        k = prob.argmax(-1)
        top2 = logits.topk(2, -1).values
        sure = (top2[:, 0] - top2[:, 1]) > 12.0
        assert torch.allclose(flow.reshape(-1, 2)[sure], offs[k][sure], atol=1e-3)


@small
@given(st.integers(0, 2**31 - 1))
def test_kalman_update_is_a_convex_combination_and_shrinks_uncertainty(seed):
    rng = np.random.default_rng(seed)
    x = torch.tensor(rng.normal(size=(1, 3, 4, 3))); z = torch.tensor(rng.normal(size=(1, 3, 4, 3)))
    su = torch.tensor(rng.uniform(1e-4, 2.0, size=(1, 3, 4, 1))); sz = torch.tensor(rng.uniform(1e-4, 2.0, size=(1, 3, 4, 1)))
    kc, ku = O.build_kf_coord(x, su, z, sz)
    lo, hi = torch.minimum(x, z), torch.maximum(x, z)
    assert bool(((kc >= lo - 1e-12) & (kc <= hi + 1e-12)).all())
    assert bool((ku <= su + 1e-15).all()) and bool((ku <= sz + 1e-15).all())
    # information form: 1/P+ = 1/P- + 1/R
    np.testing.assert_allclose((1 / ku**2).numpy(), (1 / su**2 + 1 / sz**2).numpy(), rtol=1e-9)
    assert bool((O.get_nis(z, sz, x, su) >= 0).all())


@small
@given(st.integers(0, 2**31 - 1))
def test_bilinear_sampler_reproduces_affine_images_in_the_interior(seed):
    """Away from the last row/column the sampler is ordinary bilinear interpolation: exact on affine images."""
    rng = np.random.default_rng(seed)
    H, W = 6, 7
    a, b, c = rng.normal(size=3)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=D), torch.arange(W, dtype=D), indexing="ij")
    img = (a * xs + b * ys + c)[None, ..., None]
    cx = torch.tensor(rng.uniform(0, W - 1.001, size=(1, 4, 5, 1))); cy = torch.tensor(rng.uniform(0, H - 1.001, size=(1, 4, 5, 1)))
    out = O.bilinear_sampler(img, torch.cat([cx, cy], -1))
    np.testing.assert_allclose(out.numpy(), (a * cx + b * cy + c).numpy(), atol=1e-10)


@small
@given(st.integers(0, 2**31 - 1), st.integers(1, 3))
def test_conv_same_output_shape_and_linearity(seed, stride):
    rng = np.random.default_rng(seed)
    if stride == 3:
        stride = 2
    x1, x2 = torch.tensor(rng.normal(size=(1, 6, 8, 2))), torch.tensor(rng.normal(size=(1, 6, 8, 2)))
    w = torch.tensor(rng.normal(size=(3, 3, 2, 3))); b = torch.zeros(3, dtype=D)
    y = lambda t: O.conv2d_same(t, w, b, stride, relu=False)
    assert tuple(y(x1).shape) == (1, 6 // stride, 8 // stride, 3)
    np.testing.assert_allclose(y(x1 + 2 * x2).numpy(), (y(x1) + 2 * y(x2)).numpy(), atol=1e-10)
    # conv2d_transpose is the adjoint of the stride-2 SAME conv:  <conv(x), g> == <x, deconv(g)>
    if stride == 2:
        g = torch.tensor(rng.normal(size=(1, 3, 4, 3)))
        wt = w.permute(0, 1, 2, 3)                             # [kh,kw,Cout_t=2,Cin_t=3] for the transpose
        back = O.conv2d_transpose_same(g, wt, torch.zeros(2, dtype=D), relu=False)
        np.testing.assert_allclose(float((y(x1) * g).sum()), float((x1 * back).sum()), rtol=1e-9)
