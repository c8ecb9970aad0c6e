"""class KFNet -- eager mirror of the reference's graph-building class (KFNet/KFNet.py:53-184, :303-403, :470-474).
Same constructor arguments and method names; tensors are NHWC torch CUDA tensors instead of tf.Tensors and every
method runs the corresponding libkfnet_b200 kernels immediately."""
from __future__ import annotations

import numpy as np
import torch

from .engine import Engine

_IDENTITY = np.eye(4, dtype=np.float32)


class KFNet:
    def __init__(self, images, spec, train_scoordnet=False, train_oflownet=False, dropout_rate=0.5, seed=None,
                 reuse=None, *, weights=None, engine: Engine | None = None, mode="fp32"):
        # training graphs (KFNet.py:55-63; train.py:283-285): the losses below run the fp32 training path (kfn_train_step)
        self.train_scoordnet, self.train_oflownet = bool(train_scoordnet), bool(train_oflownet)
        self.focal_x, self.focal_y, self.u, self.v = spec.focal_x, spec.focal_y, spec.u, spec.v
        self.batch_size = spec.batch_size
        self.height, self.width = spec.image_size
        self.flow_sample_rate = 8                      # KFNet.py:70
        self.min_uncertainty = 1e-5                    # KFNet.py:71
        self.engine = engine if engine is not None else Engine(weights, mode=mode, image_size=spec.image_size)
        if images.dtype != torch.uint8:                # the reference feeds decode_png output cast to float (train.py:215-217)
            images = images.round().clamp(0, 255).to(torch.uint8)
        self.images = images.to(torch.device("cuda", self.engine.device)).contiguous()
        assert self.images.shape[0] == self.batch_size
        self._measure = self.engine.scoordnet(self.images)          # BuildSCoordNet (KFNet.py:303-313)
        self.temp_feat_maps = self.BuildOFlowFeat()                 # KFNet.py:75

    # ---------------------------------------------------------------- I/O (KFNet.py:78-100)
    def GetInputImages(self):
        return self.images

    def GetMeasureCoord(self):
        return self._measure[..., :3], self._measure[..., 3:4]

    def GetMeasureCoord2(self):                                     # KFNet.py:470-474
        return self._measure[1:2, ..., :3], self._measure[1:2, ..., 3:4]

    def BuildOFlowFeat(self):                                       # KFNet.py:315-341
        return self.engine.oflow_feat(self.images)

    def BuildCoordVolume(self, feat_map1, feat_map2, window_size=8):   # KFNet.py:343-359
        assert window_size == 8
        n, h, w, c = feat_map1.shape
        diff = self.engine.diff_volume(feat_map1.contiguous(), feat_map2.contiguous())
        offs = torch.tensor([[j - 4, i - 4] for i in range(8) for j in range(8)], dtype=torch.float32, device=diff.device)
        return diff.reshape(n, h, w, 64 * c), offs

    def _step(self, feat1, feat2, last_coord, last_unc, measure):
        flow, sigw, _, _ = self.engine.oflownet(feat1.contiguous(), feat2.contiguous())
        last = torch.cat([last_coord, last_unc], -1).reshape(1, *last_coord.shape[1:3], 4).contiguous()
        return self.engine.kalman_scan(measure.unsqueeze(0).contiguous(), flow.unsqueeze(0), sigw.unsqueeze(0), _IDENTITY,
                                       state_in=last, first_index=1, reinit_period=0, diagnostics=True)

    def BuildOFlowNet(self, feat_map1, feat_map2, coord_map1, uncertainty1):   # KFNet.py:361-403
        meas = torch.cat([coord_map1, uncertainty1], -1).contiguous()          # measurement is unused by the prior
        r = self._step(feat_map1, feat_map2, coord_map1, uncertainty1, meas)
        return r["temp"][0, :, ..., :3], r["temp"][0, :, ..., 3:4]

    def GetKFCoordRecursive(self, last_coord, last_uncertainty):    # KFNet.py:85-100
        r = self._step(self.temp_feat_maps[0:1], self.temp_feat_maps[1:2], last_coord, last_uncertainty, self._measure[1:2])
        t, k = r["temp"][0], r["kf"][0]
        self._last_nis = r["nis"][0]
        return t[..., :3], t[..., 3:4], k[..., :3], k[..., 3:4]

    def GetKFCoordBatch(self):                                      # KFNet.py:102-146: frame 1 = measurement, then 3 steps
        B = self.batch_size
        flow, sigw, _, _ = self.engine.oflownet(self.temp_feat_maps[:-1].contiguous(), self.temp_feat_maps[1:].contiguous())
        z = torch.zeros(1, *flow.shape[1:], device=flow.device)
        flow = torch.cat([z, flow]).unsqueeze(0).contiguous()
        sigw = torch.cat([z[..., 0], sigw]).unsqueeze(0).contiguous()
        r = self.engine.kalman_scan(self._measure.unsqueeze(0).contiguous(), flow, sigw, _IDENTITY, first_index=0,
                                    reinit_period=1 << 30, diagnostics=True)
        t, k = r["temp"][0], r["kf"][0]
        assert t.shape[0] == B
        return t[..., :3], t[..., 3:4], k[..., :3], k[..., 3:4]

    # ---------------------------------------------------------------- losses (KFNet.py:192-298; train.py:287-296)
    def _losses(self, gt_coords, mask):
        """Forward of the training graph on this window (no parameter update): dict of the scalars train.py:`run` returns."""
        key = (gt_coords.data_ptr(), mask.data_ptr())
        if getattr(self, "_loss_cache", (None, None))[0] != key:
            if getattr(self.engine, "_train_batch", None) != self.batch_size:
                self.engine.train_begin(self.batch_size, self.train_scoordnet, self.train_oflownet)
            out = self.engine.train_step(self.images, gt_coords.float().contiguous(), mask.float().contiguous(), apply_update=False)
            self._loss_cache = (key, out)
        return self._loss_cache[1]

    def MeasureCoordLoss(self, gt_coords, mask, transform=None, pixel_map=None, poses=None, images=None):   # KFNet.py:234-255
        assert transform is None and pixel_map is None and poses is None, "train.py passes images= only (train.py:287-292)"
        o = self._losses(gt_coords, mask)
        return o["measure_loss"], o["measure_accuracy"]

    def TemporalCoordLoss(self, gt_coords, mask, transform=None, pixel_map=None, poses=None, images=None):  # KFNet.py:257-279
        assert transform is None and pixel_map is None and poses is None
        o = self._losses(gt_coords, mask)
        return o["temporal_loss"], o["temporal_accuracy"]

    def KFCoordLoss(self, gt_coords, mask, transform=None, pixel_map=None, poses=None, images=None):        # KFNet.py:281-300
        assert transform is None and pixel_map is None and poses is None
        o = self._losses(gt_coords, mask)
        return o["kf_loss"], o["kf_accuracy"]

    # ---------------------------------------------------------------- pointwise API surface (kfn_kalman_update)
    def BuildKFCoord(self, last_coord, last_uncertainty, measure_coord, measure_uncertainty):   # KFNet.py:148-162
        prior = torch.cat([last_coord, last_uncertainty], -1).float().contiguous()
        meas = torch.cat([measure_coord, measure_uncertainty], -1).float().contiguous()
        post, _ = self.engine.kalman_update(prior, meas)
        return post[..., :3], post[..., 3:4]

    def GetNIS(self, measure_coord_map, measure_uncertainty_map, temp_coord_map, temp_uncertainty_map):   # KFNet.py:174-184
        prior = torch.cat([temp_coord_map, temp_uncertainty_map], -1).float().contiguous()
        meas = torch.cat([measure_coord_map, measure_uncertainty_map], -1).float().contiguous()
        _, nis = self.engine.kalman_update(prior, meas, want_nis=True)
        return nis
