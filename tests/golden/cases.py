"""Seeded inputs shared by make_golden.py (runs the reference source on the TF1 shim, in the build
container only) and by the tests that compare the oracle / CUDA path with the committed fixtures."""
import numpy as np

from kfnet_b200 import synth

SMALL_HW = (64, 96)          # image size of the small cases -> 8x12 label grid
WEIGHT_SEED = 0


def weights():
    return synth.make_weights(WEIGHT_SEED)


def small_images(n, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, size=(SMALL_HW[0] + 16, SMALL_HW[1] + 16, 3)).astype(np.float32)
    # blur a little so neighbouring frames correlate, then shift per frame
    k = np.ones((3, 3), np.float32) / 9.0
    sm = np.zeros_like(base)
    for dy in range(3):
        for dx in range(3):
            sm += k[dy, dx] * np.roll(np.roll(base, dy - 1, 0), dx - 1, 1)
    out = np.empty((n,) + SMALL_HW + (3,), np.float32)
    for i in range(n):
        oy, ox = int(rng.integers(0, 16)), int(rng.integers(0, 16))
        out[i] = np.floor(sm[oy:oy + SMALL_HW[0], ox:ox + SMALL_HW[1]])
    return out


def small_state(seed):
    rng = np.random.default_rng(seed)
    h, w = SMALL_HW[0] // 8, SMALL_HW[1] // 8
    coord = rng.normal(size=(1, h, w, 3))
    unc = rng.uniform(0.02, 0.3, size=(1, h, w, 1))
    return coord, unc


def bilinear_case():
    """imgs [2,5,7,3]; coords cover interior, exact integers, x=W-1, x in [W-1,W), negatives, far OOB."""
    rng = np.random.default_rng(11)
    imgs = rng.normal(size=(2, 5, 7, 3))
    xs = np.array([0.0, 0.25, 3.5, 5.0, 5.75, 6.0, 6.3, 6.999, 7.0, 9.5, -0.5, -1.0, -2.5, 2.0])
    ys = np.array([0.0, 1.5, 3.0, 3.9, 4.0, 4.2, 5.0, -0.3, -1.7, 2.0])
    gx, gy = np.meshgrid(xs, ys)
    coords = np.stack([gx, gy], -1)[None].repeat(2, 0)          # [2,10,14,2]
    return imgs, coords


def volume_case():
    rng = np.random.default_rng(12)
    f1 = rng.normal(size=(1, 9, 11, 4))
    f2 = rng.normal(size=(1, 9, 11, 4))
    return f1, f2


def full_images():
    return synth.make_sequence(2, seed=1).astype(np.float32)


def full_state():
    rng = np.random.default_rng(21)
    return rng.normal(size=(1, 60, 80, 3)), rng.uniform(0.02, 0.3, size=(1, 60, 80, 1))


def transform():
    return np.linalg.inv(synth.make_transform(2)).astype(np.float64)



def label_case(n=1, seed=31):
    """Full-resolution labels for the printed metrics (train.py:219-222, KFNet.py:192-232, eval.py:10-29):
    gt coords [n,H,W,3], mask [n,H,W,1] with 1.0 / 0.0 and a few other values (the reference keeps only mask == 1.0
    after the nearest-neighbour downsample), predictions near the ground truth so both sides of the 5 cm accuracy
    threshold and of the loss clamp occur, sigma spanning the min_uncertainty clamp."""
    rng = np.random.default_rng(seed)
    H, W = SMALL_HW
    gt = rng.normal(size=(n, H, W, 3))
    mask = (rng.uniform(size=(n, H, W, 1)) > 0.3).astype(np.float64)
    mask[:, ::5, ::7] = 0.5                                       # neither 0 nor 1
    h, w = H // 8, W // 8
    gt8 = gt[:, ::8, ::8]
    scale = np.where(rng.uniform(size=(n, h, w, 1)) > 0.5, 0.01, 0.2)
    pred_world = gt8 + rng.normal(size=(n, h, w, 3)) * scale     # what T . pred must equal
    T = transform()
    pred = (pred_world - T[:3, 3]) @ np.linalg.inv(T[:3, :3]).T   # prediction in the network's frame
    unc = np.exp(rng.uniform(np.log(3e-6), np.log(0.5), size=(n, h, w, 1)))
    nis = np.exp(rng.normal(size=(1, h, w, 3)) * 2.0 - 1.0) * (rng.uniform(size=(1, h, w, 3)) > 0.1)
    return dict(gt=gt, mask=mask, pred=pred, unc=unc, nis=nis)


def training_case(seed=41):
    """4 consecutive small frames; ground truth = the (oracle-computed) measurement map of the synthetic network plus
    noise of two scales, moved to the world frame -- so that both branches of the loss clamp (KFNet.py:217: only
    pixels whose unclamped loss is below -2 carry gradient) and of the 5 cm accuracy threshold occur; mask has zeros;
    poses / pixel map for ReprojectionLoss.  (The oracle only shapes the INPUTS here; the golden outputs come from the
    reference's code.)"""
    import torch
    from oracle import kfnet_oracle as O
    rng = np.random.default_rng(seed)
    imgs = small_images(4, seed=seed)
    h, w = SMALL_HW[0] // 8, SMALL_HW[1] // 8
    pred = O.scoordnet(imgs, weights(), torch.float64)[0].numpy()
    scale = np.where(rng.uniform(size=(4, h, w, 1)) > 0.4, 0.004, 0.3)
    gt_net = pred + rng.normal(size=pred.shape) * scale
    T = transform()
    gt_world = gt_net @ T[:3, :3].T + T[:3, 3]
    mask = (rng.uniform(size=(4, h, w, 1)) > 0.25).astype(np.float64)
    poses = np.stack([np.linalg.inv(T) for _ in range(4)])
    poses[:, :3, 3] += rng.normal(size=(4, 3)) * 0.1
    ys, xs = np.meshgrid(np.linspace(-0.4, 0.4, h), np.linspace(-0.5, 0.5, w), indexing="ij")
    pixel_map = np.stack([xs, ys], -1)[None].repeat(4, 0)
    return dict(images=imgs, gt_world=gt_world, mask=mask, poses=poses, pixel_map=pixel_map)
