"""Name-for-name mirrors of the reference's small helpers on the path, running on the C ABI:
  tools/util.py   : bilinear_sampler
  KFNet/util.py   : HomoCoord, ApplyTransform, GetPixelMap
(the hot path uses the fused forms inside kfn_kalman_scan; these exist for the call surface and for direct tests)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def bilinear_sampler(imgs, coords):
    """tools/util.py:3-94.  imgs [N,H,W,C], coords [N,Ht,Wt,2] (x,y) -> [N,Ht,Wt,C] (fp32 CUDA tensors)."""
    imgs, coords = imgs.float().contiguous(), coords.float().contiguous()
    N, H, W, Cc = imgs.shape
    _, Ht, Wt, two = coords.shape
    assert two == 2 and coords.shape[0] == N and imgs.is_cuda and coords.is_cuda
    out = torch.empty(N, Ht, Wt, Cc, dtype=torch.float32, device=imgs.device)
    _lib.check(_lib.load().kfn_bilinear_sampler(C.c_void_p(imgs.data_ptr()), C.c_void_p(coords.data_ptr()),
                                                C.c_void_p(out.data_ptr()), N, H, W, Cc, Ht, Wt, _stream(imgs)))
    return out


def HomoCoord(coords):
    """KFNet/util.py:5-10 (pure data movement)."""
    return torch.cat([coords, torch.ones_like(coords[..., :1])], -1)


def ApplyTransform(coords, transform, inverse=False):
    """KFNet/util.py:12-40.  coords [...,3] CUDA fp32; transform 4x4 (numpy / tensor); inverse -> matrix_inverse first."""
    T = np.asarray(transform.detach().cpu() if isinstance(transform, torch.Tensor) else transform, dtype=np.float32).reshape(4, 4)
    if inverse:
        T = np.linalg.inv(T).astype(np.float32)
    x = coords.float().contiguous()
    out = torch.empty_like(x)
    T16 = np.ascontiguousarray(T.reshape(16))
    _lib.check(_lib.load().kfn_apply_transform(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), x.numel() // 3,
                                               T16.ctypes.data_as(C.c_void_p), _stream(x)))
    return out


def GetPixelMap(batch_size, height, width, normalize=False, spec=None, device="cuda"):
    """KFNet/util.py:42-63: [B,H,W,2] with [...,0] = x, [...,1] = y (integer grid; normalised by the intrinsics if asked)."""
    ys, xs = torch.meshgrid(torch.arange(height, dtype=torch.float32), torch.arange(width, dtype=torch.float32), indexing="ij")
    if normalize:
        xs, ys = (xs - spec.u) / spec.focal_x, (ys - spec.v) / spec.focal_y
    return torch.stack([xs, ys], -1)[None].repeat(batch_size, 1, 1, 1).to(device)
