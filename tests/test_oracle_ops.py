"""Op-level pins for the oracle: naive numpy loops written from the TF-1.13 op definitions
 vs the oracle's torch implementations, plus hand-computable micro cases."""
import numpy as np
import pytest
import torch

from oracle import kfnet_oracle as O

D = torch.float64


def naive_conv_same(x, w, b, s):
    n, h, wd, cin = x.shape
    kh, kw, _, cout = w.shape
    oh, ow = -(-h // s), -(-wd // s)
    pt = max((oh - 1) * s + kh - h, 0) // 2
    pl = max((ow - 1) * s + kw - wd, 0) // 2
    y = np.zeros((n, oh, ow, cout))
    for i in range(oh):
        for j in range(ow):
            for r in range(kh):
                for c in range(kw):
                    yy, xx = i * s + r - pt, j * s + c - pl
                    if 0 <= yy < h and 0 <= xx < wd:
                        y[:, i, j, :] += x[:, yy, xx, :] @ w[r, c]
    return y + b


def naive_deconv_same(x, w, b):
    """out[j] = sum_{i,k: 2i+k=j} in[i] w[k], j in [0,2n): scatter form of App. A.3."""
    n, h, wd, cin = x.shape
    kh, kw, cout, _ = w.shape
    y = np.zeros((n, 2 * h, 2 * wd, cout))
    for i in range(h):
        for j in range(wd):
            for r in range(kh):
                for c in range(kw):
                    yy, xx = 2 * i + r, 2 * j + c
                    if yy < 2 * h and xx < 2 * wd:
                        y[:, yy, xx, :] += x[:, i, j, :] @ w[r, c].T
    return y + b


@pytest.mark.parametrize("h,w,k,s", [(6, 8, 3, 1), (6, 8, 3, 2), (8, 8, 1, 1), (2, 2, 3, 2), (1, 1, 3, 1), (4, 4, 3, 2)])
def test_conv_same_matches_naive(h, w, k, s):
    rng = np.random.default_rng(0)
    x = rng.normal(size=(2, h, w, 3)); wt = rng.normal(size=(k, k, 3, 5)); b = rng.normal(size=5)
    got = O.conv2d_same(torch.tensor(x), torch.tensor(wt), torch.tensor(b), s, relu=False).numpy()
    np.testing.assert_allclose(got, naive_conv_same(x, wt, b, s), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("h,w,k,s", [(6, 8, 3, 1), (6, 8, 3, 2), (7, 9, 3, 2), (8, 12, 1, 1), (5, 5, 3, 1), (64, 96, 3, 2)])
def test_conv_same_matches_third_party_tf_padding(h, w, k, s):
    """Independent statement of TensorFlow's SAME padding that ships in this image: Hugging Face's `apply_tf_padding`
    (written to run TF-trained MobileNet checkpoints in PyTorch, citing tf.nn's padding notes) + a VALID torch conv."""
    hf = pytest.importorskip("transformers.models.mobilenet_v1.modeling_mobilenet_v1")
    rng = np.random.default_rng(1)
    x = rng.normal(size=(2, h, w, 3)); wt = rng.normal(size=(k, k, 3, 5)); b = rng.normal(size=5)
    conv = torch.nn.Conv2d(3, 5, k, stride=s, padding=0, bias=True).double()
    with torch.no_grad():
        conv.weight.copy_(torch.tensor(wt).permute(3, 2, 0, 1)); conv.bias.copy_(torch.tensor(b))
        ref = conv(hf.apply_tf_padding(torch.tensor(x).permute(0, 3, 1, 2), conv)).permute(0, 2, 3, 1).numpy()
    got = O.conv2d_same(torch.tensor(x), torch.tensor(wt), torch.tensor(b), s, relu=False).numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12)


def test_conv_same_stride2_pads_bottom_right_only():
    x = np.zeros((1, 4, 4, 1)); x[0, 0, 0, 0] = 1.0
    w = np.zeros((3, 3, 1, 1)); w[0, 0, 0, 0] = 1.0           # tap (0,0) must sit on input (2i,2j)
    y = O.conv2d_same(torch.tensor(x), torch.tensor(w), torch.zeros(1, dtype=D), 2, relu=False).numpy()
    assert y[0, 0, 0, 0] == 1.0 and y.sum() == 1.0


@pytest.mark.parametrize("n", [1, 2, 4])
def test_deconv_same_matches_naive(n):
    rng = np.random.default_rng(1)
    x = rng.normal(size=(2, n, n, 4)); wt = rng.normal(size=(3, 3, 6, 4)); b = rng.normal(size=6)
    got = O.conv2d_transpose_same(torch.tensor(x), torch.tensor(wt), torch.tensor(b), relu=False).numpy()
    np.testing.assert_allclose(got, naive_deconv_same(x, wt, b), rtol=1e-12, atol=1e-12)


def test_translate_direction_and_zero_fill():
    img = torch.arange(12, dtype=D).reshape(1, 3, 4, 1) + 1
    out = O.translate_nearest(img, 1, 0)          # out[y,x] = img[y, x-1]
    assert out[0, :, 0, 0].abs().sum() == 0 and torch.equal(out[0, :, 1:, 0], img[0, :, :-1, 0])
    out = O.translate_nearest(img, 0, -1)         # out[y,x] = img[y+1, x]
    assert out[0, 2, :, 0].abs().sum() == 0 and torch.equal(out[0, :2, :, 0], img[0, 1:, :, 0])


def test_coord_volume_semantics():
    """diff[p, k] = f2[p] - f1[p + (ox,oy)], zero-filled shift => border diff == f2 (KFNet.py:351-353)."""
    rng = np.random.default_rng(2)
    f1 = torch.tensor(rng.normal(size=(1, 6, 7, 2))); f2 = torch.tensor(rng.normal(size=(1, 6, 7, 2)))
    diff, offs = O.coord_volume(f1, f2, 8)
    assert offs[0].tolist() == [-4, -4] and offs[1].tolist() == [-3, -4] and offs[63].tolist() == [3, 3]
    d = diff.reshape(1, 6, 7, 64, 2)
    for k in (0, 9, 36, 63):
        ox, oy = int(offs[k, 0]), int(offs[k, 1])
        for y in range(6):
            for x in range(7):
                yy, xx = y + oy, x + ox
                ref = f2[0, y, x] - (f1[0, yy, xx] if 0 <= yy < 6 and 0 <= xx < 7 else 0)
                assert torch.equal(d[0, y, x, k], ref)


def test_bilinear_edge_quirk():
    """App. A.6: for x in [W-1, W) both taps hit column W-1 with weights -a,+a => ~0 contribution."""
    img = torch.ones(1, 2, 4, 1, dtype=D) * 5.0
    co = lambda x, y: torch.tensor([[[[x, y]]]], dtype=D)
    assert abs(O.bilinear_sampler(img, co(1.5, 0.5)).item() - 5.0) < 1e-12
    assert abs(O.bilinear_sampler(img, co(3.0, 0.0)).item()) < 1e-12       # exactly x = W-1 -> 0
    assert abs(O.bilinear_sampler(img, co(3.3, 0.0)).item()) < 1e-12
    assert abs(O.bilinear_sampler(img, co(-0.5, 0.0)).item()) < 1e-12
    _, taps = O.bilinear_sampler(img, co(2.7, -0.2), return_taps=True)
    assert taps.reshape(-1).tolist() == [2, -1]


def test_kalman_update_and_nis():
    x = torch.tensor([[[[1.0, 2.0, 3.0]]]], dtype=D); su = torch.tensor([[[[0.3]]]], dtype=D)
    z = torch.tensor([[[[2.0, 0.0, 3.0]]]], dtype=D); sz = torch.tensor([[[[0.1]]]], dtype=D)
    kc, ku = O.build_kf_coord(x, su, z, sz)
    K = 0.09 / (0.09 + 0.01)
    np.testing.assert_allclose(kc.numpy().ravel(), (1 - K) * x.numpy().ravel() + K * z.numpy().ravel(), rtol=1e-14)
    np.testing.assert_allclose(ku.item(), np.sqrt((1 - K) * 0.09), rtol=1e-14)
    assert ku.item() <= su.item()
    nis = O.get_nis(z, sz, x, su)
    np.testing.assert_allclose(nis.numpy().ravel(), np.array([1.0, 4.0, 0.0]) / 0.1, rtol=1e-12)


def test_flow_range_and_softargmax():
    rng = np.random.default_rng(3)
    prob = torch.softmax(torch.tensor(rng.normal(size=(10, 64)) * 5), -1)
    _, offs = O.coord_volume(torch.zeros(1, 1, 1, 1, dtype=D), torch.zeros(1, 1, 1, 1, dtype=D), 8)
    flow, pm = O.flow_from_prob(prob, offs, 1, 2, 5)
    assert flow.min() >= -4 and flow.max() <= 3
    onehot = torch.zeros(1, 64, dtype=D); onehot[0, 5 * 8 + 2] = 1           # (i=5,j=2) -> (ox,oy)=(-2,1)
    f, _ = O.flow_from_prob(onehot, offs, 1, 1, 1)
    assert f.reshape(-1).tolist() == [-2.0, 1.0]
