"""Executable model of the index arithmetic of kfnet_b200/csrc/conv1ab_fused.cu (the opt-in fused conv1a+conv1b kernel).

The kernel could not be run before the end of round 1, so its data movement is restated here row for row -- halo-pixel
numbering of the producers, the two 128-row MMA-1 tiles, epilogue-1's zeroing rule (= conv1b's SAME padding) and
interior test (feat1), MMA-2's shifted 8-row groups pitched by one haloed image row, epilogue-2's row -> pixel map --
and checked against the oracle's conv1b(relu(conv1a(x))) on images whose tiles touch every border.  Swizzles, barriers
and TMEM are not modelled (the GPU test tests/test_gpu_conv_layers.py::test_fused_conv1ab covers the real thing)."""
import numpy as np
import pytest
import torch

from oracle import kfnet_oracle as O

HALO_W, HALO_H, HALO_ROWS = 10, 18, 180          # kHaloW, kHaloH, kHaloRows


def fused_model(img, w1, b1, w2, b2, wf=None, bf=None):
    """img u8-valued float [N,H,W,3]; w1 [3,3,3,64] (already includes the 0.00625 scale), w2 [3,3,64,64] HWIO."""
    N, H, W, _ = img.shape
    assert H % 16 == 0 and W % 8 == 0
    out = np.zeros((N, H, W, 64))
    feat1 = np.full((N, H, W, wf.shape[-1]), np.nan) if wf is not None else None
    B1 = w1.reshape(27, 64)                                        # k = (ky*3 + kx)*3 + c
    BF = wf.reshape(27, -1) if wf is not None else None
    tap_off = [(t // 3) * HALO_W + (t % 3) for t in range(9)]      # tap_off16 in rows: (ky, kx) = (1+dy, 1+dx)
    for n in range(N):
        for th in range(H // 16):
            for tw in range(W // 8):
                # ---- producers: A1 row r = halo pixel (hy, hx), 32-wide im2col row (27 used)
                A1 = np.zeros((256, 27))
                for r in range(192):
                    hy, hx = divmod(r, HALO_W)
                    y, x = th * 16 - 1 + hy, tw * 8 - 1 + hx
                    inside = r < HALO_ROWS and 0 <= y < H and 0 <= x < W
                    if not inside:
                        continue
                    for ky in range(3):
                        yy = y + ky - 1
                        for jj in range(9):
                            xx = x - 1 + jj // 3
                            if 0 <= yy < H and 0 <= xx < W:
                                A1[r, ky * 9 + jj] = img[n, yy, xx, jj % 3] - 128.0
                # ---- MMA-1 (two 128-row tiles) + epilogue-1
                acc1 = A1 @ B1                                      # rows 0..255
                A2 = np.zeros((HALO_ROWS, 64))
                for r in range(HALO_ROWS):
                    hy, hx = divmod(r, HALO_W)
                    y, x = th * 16 - 1 + hy, tw * 8 - 1 + hx
                    inside = 0 <= y < H and 0 <= x < W
                    A2[r] = np.maximum(acc1[r] + b1, 0.0) if inside else 0.0
                    if feat1 is not None and 1 <= hy <= 16 and 1 <= hx <= 8:
                        feat1[n, y, x] = np.maximum(A1[r] @ BF + bf, 0.0)
                # ---- MMA-2: output row m = h*8 + w reads A2 row tap_off + (m // 8) * 10 + (m % 8)
                acc2 = np.zeros((128, 64))
                for t in range(9):
                    ky, kx = divmod(t, 3)
                    rows = [tap_off[t] + (m // 8) * HALO_W + (m % 8) for m in range(128)]
                    acc2 += A2[rows] @ w2[ky, kx]
                # ---- epilogue-2: row -> pixel
                for m in range(128):
                    rh, rw = m >> 3, m & 7
                    out[n, th * 16 + rh, tw * 8 + rw] = np.maximum(acc2[m] + b2, 0.0)
    return out, feat1


@pytest.mark.parametrize("shape", [(1, 16, 8), (2, 32, 24), (1, 48, 16)])
def test_fused_conv1ab_index_model_matches_oracle(shape):
    n, h, w = shape
    rng = np.random.default_rng(3 + h)
    img = rng.integers(0, 256, size=(n, h, w, 3)).astype(np.float64)
    w1 = rng.normal(size=(3, 3, 3, 64)) * 0.01; b1 = rng.normal(size=64) * 0.1
    wf = rng.normal(size=(3, 3, 3, 16)) * 0.01; bf = rng.normal(size=16) * 0.1
    w2 = rng.normal(size=(3, 3, 64, 64)) * 0.05; b2 = rng.normal(size=64) * 0.1
    got, f1 = fused_model(img, w1, b1, w2, b2, wf, bf)
    x = torch.tensor(img - 128.0)
    mid = O.conv2d_same(x, torch.tensor(w1), torch.tensor(b1), 1, True)
    ref = O.conv2d_same(mid, torch.tensor(w2), torch.tensor(b2), 1, True).numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-10)
    reff = O.conv2d_same(x, torch.tensor(wf), torch.tensor(bf), 1, True).numpy()
    assert not np.isnan(f1).any()                                  # every pixel is some tile's interior exactly once
    np.testing.assert_allclose(f1, reff, rtol=1e-10, atol=1e-10)


# ------------------------------------------------------------------------------------------------------------------
# conv_tc_kernel_diff (KFN_TC_DIFF_FUSED=1, opt-in): U-Net conv0 whose producer warps form the difference volume.
def diff_conv0_model(f_prev, f_cur, w, b):
    """f_prev/f_cur [NP,GH,GW,C]; w [3,3,C,Cout] HWIO.  Mirrors the producer's row numbering (halo tile of two 8x8
    windows, rows [y][n][x], ring = zeros), the (i, j) -> (oy, ox) = (i-4, j-4) shift with zero fill, and the halo MMA
    path's tap offsets / 8-row groups pitched by 10*hbn rows."""
    NP, GH, GW, C = f_cur.shape
    P, N = GH * GW, NP * GH * GW
    cout = w.shape[-1]
    out = np.zeros((N, 8, 8, cout))
    fp, fc = f_prev.reshape(NP, P, C), f_cur.reshape(N, C)
    hbn = 2
    tap_off = [(1 + dy) * 10 * hbn + (1 + dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    for tile in range((N + 1) // 2):
        A = np.zeros((224, C))
        for r in range(224):
            if r >= 200:
                continue
            y, rem = divmod(r, 20)
            wn, x = divmod(rem, 10)
            window = tile * 2 + wn
            ring = not (1 <= y <= 8 and 1 <= x <= 8)
            if ring or window >= N:
                continue
            pair, pix = divmod(window, P)
            py, px = divmod(pix, GW)
            sy, sx = py + (y - 1) - 4, px + (x - 1) - 4
            prev = fp[pair, sy * GW + sx] if (0 <= sy < GH and 0 <= sx < GW) else 0.0
            A[r] = fc[window] - prev
        acc = np.zeros((128, cout))                     # output row m = (h, n, w): m = h*16 + n*8 + w
        for t in range(9):
            ky, kx = divmod(t, 3)
            rows = [tap_off[t] + (m // 8) * 10 + (m % 8) for m in range(128)]   # 8-row groups pitched by 10 rows
            acc += A[rows] @ w[ky, kx]
        for m in range(128):
            hh, rest = divmod(m, 16)
            wn, ww = divmod(rest, 8)
            window = tile * 2 + wn
            if window < N:
                out[window, hh, ww] = np.maximum(acc[m] + b, 0.0)
    return out


def test_fused_diff_conv0_index_model_matches_oracle():
    rng = np.random.default_rng(9)
    NP, GH, GW, C, cout = 1, 5, 7, 8, 6                 # odd window count: the last tile is half empty
    f_prev, f_cur = rng.normal(size=(NP, GH, GW, C)), rng.normal(size=(NP, GH, GW, C))
    w, b = rng.normal(size=(3, 3, C, cout)) * 0.2, rng.normal(size=cout) * 0.1
    got = diff_conv0_model(f_prev, f_cur, w, b)
    vol, _ = O.coord_volume(torch.tensor(f_prev), torch.tensor(f_cur))          # [NP,GH,GW,64*C]
    x = vol.reshape(NP * GH * GW, 8, 8, C)                                      # KFNet.py:372
    ref = O.conv2d_same(x, torch.tensor(w), torch.tensor(b), 1, True).numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-10)
