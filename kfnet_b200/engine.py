"""Engine -- thin Python owner of one kfn_handle: weights in, device tensors through the C ABI.
torch tensors are only the allocation / stream carrier; every kernel is in libkfnet_b200.so."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .synth import weight_shapes

MODES = {"fp32": _lib.KFN_MODE_FP32, "bf16": _lib.KFN_MODE_BF16, "bf16x3": _lib.KFN_MODE_BF16X3, "fp16x2": _lib.KFN_MODE_FP16X2}


def _ptr(t):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


class Engine:
    def __init__(self, weights: dict, mode: str = "fp32", image_size=(480, 640), device: int | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("kfnet_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.mode = mode
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.H, self.W = int(image_size[0]), int(image_size[1])
        self.GH, self.GW = self.H // 8, self.W // 8
        h = C.c_void_p()
        _lib.check(self.lib.kfn_create(C.byref(h), self.device, MODES[mode], self.H, self.W))
        self.h = h
        self._ws = None
        self._graphs = {}
        self.load_weights(weights)

    # ---- weights (replaces tf.train.Saver restore by variable name, train.py:317-321) ----
    def load_weights(self, weights: dict):
        for name in weight_shapes():
            if name not in weights:
                raise KeyError(f"missing variable {name}")
            a = np.ascontiguousarray(np.asarray(weights[name], dtype=np.float32))
            dims = (C.c_int64 * a.ndim)(*a.shape)
            _lib.check(self.lib.kfn_load_weight(self.h, name.encode(), a.ctypes.data_as(C.c_void_p), dims, a.ndim))
        _lib.check(self.lib.kfn_finalize_weights(self.h))
        self._graphs = {}          # packed weights were re-allocated: graphs captured before are stale

    def close(self):
        if getattr(self, "h", None):
            self.lib.kfn_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers ----
    def _dev(self):
        return torch.device("cuda", self.device)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _workspace(self, nbytes: int):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._graphs = {}      # captured graphs have the old workspace pointer baked in: drop them with it
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self._dev())
        return self._ws

    def _chk_img(self, images):
        assert images.is_cuda and images.dtype == torch.uint8 and images.is_contiguous()
        assert tuple(images.shape[-3:]) == (self.H, self.W, 3), images.shape

    def launch_count(self, reset=False) -> int:
        return int(self.lib.kfn_launch_count(self.h, 1 if reset else 0))

    # ---- SCoordNet: images u8 [B,H,W,3] -> measure fp32 [B,GH,GW,4] = (xyz, sigma_z) ----
    def scoordnet(self, images):
        self._chk_img(images)
        B = images.shape[0]
        out = torch.empty(B, self.GH, self.GW, 4, dtype=torch.float32, device=self._dev())
        ws = self._workspace(self.lib.kfn_workspace_bytes(self.h, B, 0))
        _lib.check(self.lib.kfn_scoordnet_forward(self.h, _ptr(images), B, _ptr(out), _ptr(ws), ws.numel(), self._stream()))
        return out

    # ---- end to end from HOST memory through the C ABI (kfn_eval_sequences_host): pinned uint8 frames in, pinned maps out ----
    def eval_sequences_host(self, host_images, transform, host_out=None, reinit_period=500, nis_gate=False,
                            frames_per_batch=8, pairs_per_batch=4):
        assert host_images.dim() == 5 and host_images.dtype == torch.uint8 and not host_images.is_cuda and host_images.is_contiguous()
        S, T = host_images.shape[:2]
        assert tuple(host_images.shape[-3:]) == (self.H, self.W, 3)
        if host_out is None:
            host_out = torch.empty(S, T, self.GH, self.GW, 4, dtype=torch.float32).pin_memory()
        key = (S, T)
        if getattr(self, "_host_stage_key", None) != key:
            self._host_stage = (torch.empty(S, T, self.H, self.W, 3, dtype=torch.uint8, device=self._dev()),
                                torch.empty(S, T, self.GH, self.GW, 4, dtype=torch.float32, device=self._dev()))
            self._host_stage_key = key
        dev_imgs, dev_out = self._host_stage
        Tm = np.ascontiguousarray(np.asarray(transform, dtype=np.float32).reshape(16))
        need = (self.lib.kfn_workspace_bytes(self.h, frames_per_batch, pairs_per_batch)
                + self.lib.kfn_sequence_extra_bytes(self.h, S, T))
        ws = self._workspace(need)
        _lib.check(self.lib.kfn_eval_sequences_host(self.h, C.c_void_p(host_images.data_ptr()), S, T, Tm.ctypes.data_as(C.c_void_p),
                                                    int(reinit_period), 1 if nis_gate else 0, int(frames_per_batch),
                                                    int(pairs_per_batch), C.c_void_p(host_out.data_ptr()), _ptr(dev_imgs),
                                                    _ptr(dev_out), _ptr(ws), ws.numel(), self._stream()))
        return host_out

    # ---- streaming form for one sequence delivered in chunks (kfn_eval_stream_host; what the drop-in eval() calls) ----
    def eval_stream_host(self, host_images, first_index, transform, carry=None, reinit_period=500, nis_gate=False,
                         frames_per_batch=16, pairs_per_batch=16, want_diag=False, capacity=None):
        """host_images: pinned uint8 [T,H,W,3] = frames first_index .. first_index+T-1 of the stream.  `carry` (device fp32
        [GH*GW*36]: filter state + last feature map) links consecutive chunks; pass the tensor returned by the previous call.
        Returns (out_map [T,GH,GW,4] pinned host view, diag dict of host views or None, carry)."""
        assert host_images.dim() == 4 and host_images.dtype == torch.uint8 and not host_images.is_cuda and host_images.is_contiguous()
        T = host_images.shape[0]
        assert tuple(host_images.shape[1:]) == (self.H, self.W, 3)
        cap = max(int(capacity or 0), T)
        P = self.GH * self.GW
        st = getattr(self, "_stream_stage", None)
        if st is None or st["cap"] < cap:
            dev = self._dev()
            st = dict(cap=cap,
                      dev_imgs=torch.empty(cap, self.H, self.W, 3, dtype=torch.uint8, device=dev),
                      dev_out=torch.empty(cap * P * 4, dtype=torch.float32, device=dev),
                      dev_diag=torch.empty(cap * P * 15, dtype=torch.float32, device=dev),
                      host_out=torch.empty(cap * P * 4, dtype=torch.float32).pin_memory(),
                      host_diag=torch.empty(cap * P * 15, dtype=torch.float32).pin_memory())
            self._stream_stage = st
        if carry is None:
            carry = torch.zeros(P * 36, dtype=torch.float32, device=self._dev())
        Tm = np.ascontiguousarray(np.asarray(transform, dtype=np.float32).reshape(16))
        need = (self.lib.kfn_workspace_bytes(self.h, frames_per_batch, pairs_per_batch)
                + self.lib.kfn_sequence_extra_bytes(self.h, 1, cap))
        ws = self._workspace(need)
        _lib.check(self.lib.kfn_eval_stream_host(
            self.h, C.c_void_p(host_images.data_ptr()), T, int(first_index), Tm.ctypes.data_as(C.c_void_p), int(reinit_period),
            1 if nis_gate else 0, int(frames_per_batch), int(pairs_per_batch), _ptr(carry),
            C.c_void_p(st["host_out"].data_ptr()), C.c_void_p(st["host_diag"].data_ptr()) if want_diag else C.c_void_p(0),
            _ptr(st["dev_imgs"]), _ptr(st["dev_out"]), _ptr(st["dev_diag"]) if want_diag else C.c_void_p(0),
            _ptr(ws), ws.numel(), self._stream()))
        out = st["host_out"][:T * P * 4].view(T, self.GH, self.GW, 4)
        diag = None
        if want_diag:
            d = st["host_diag"]
            diag = dict(measure=d[:T * P * 4].view(T, self.GH, self.GW, 4),
                        temp=d[T * P * 4:T * P * 8].view(T, self.GH, self.GW, 4),
                        kf=d[T * P * 8:T * P * 12].view(T, self.GH, self.GW, 4),
                        nis=d[T * P * 12:T * P * 15].view(T, self.GH, self.GW, 3))
        return out, diag, carry

    # ---- the same, replayed from a CUDA graph (captured on first use for a given set of buffers / parameters):
    #      removes the CPU launch cost and most of the inter-kernel gaps of the ~7000 launches of a 256-frame step ----
    def eval_sequences_graphed(self, images, transform, out, reinit_period=500, nis_gate=False, frames_per_batch=8,
                               pairs_per_batch=4):
        key = (images.data_ptr(), out.data_ptr(), tuple(images.shape), int(reinit_period), bool(nis_gate),
               int(frames_per_batch), int(pairs_per_batch), np.asarray(transform, dtype=np.float32).tobytes())
        g = self._graphs.get(key)
        if g is None:
            # warm-up outside capture (workspace allocation, TMA descriptor cache, function attributes)
            self.eval_sequences(images, transform, reinit_period, nis_gate, frames_per_batch, pairs_per_batch, out=out)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.eval_sequences(images, transform, reinit_period, nis_gate, frames_per_batch, pairs_per_batch, out=out)
            if len(self._graphs) > 8:
                self._graphs.clear()
            self._graphs[key] = g
        g.replay()
        return out

    # ---- feature tower: images -> l2-normalised feat fp32 [B,GH,GW,32] ----
    def oflow_feat(self, images):
        self._chk_img(images)
        B = images.shape[0]
        out = torch.empty(B, self.GH, self.GW, 32, dtype=torch.float32, device=self._dev())
        ws = self._workspace(self.lib.kfn_workspace_bytes(self.h, B, 0))
        _lib.check(self.lib.kfn_oflow_feat_forward(self.h, _ptr(images), B, _ptr(out), _ptr(ws), ws.numel(), self._stream()))
        return out

    def diff_volume(self, f_prev, f_cur, bf16=False, out=None):
        NP, GH, GW, Cc = f_prev.shape
        assert Cc == 32 and f_prev.dtype == torch.float32 and f_prev.is_contiguous() and f_cur.is_contiguous()
        if out is None:
            out = torch.empty(NP * GH * GW, 8, 8, 32, dtype=torch.bfloat16 if bf16 else torch.float32, device=f_prev.device)
        _lib.check(self.lib.kfn_diff_volume(_ptr(f_prev), _ptr(f_cur), NP, GH, GW, _ptr(out), 1 if bf16 else 0, self._stream()))
        return out

    # ---- OFlowNet on NP pairs -> flow [NP,GH,GW,2], sigma_w [NP,GH,GW] (+ logits [NP*P,64], argmax) ----
    def oflownet(self, f_prev, f_cur, want_logits=False, want_argmax=False):
        assert f_prev.shape == f_cur.shape and f_prev.is_contiguous() and f_cur.is_contiguous()
        NP = f_prev.shape[0]
        dev = f_prev.device
        flow = torch.empty(NP, self.GH, self.GW, 2, dtype=torch.float32, device=dev)
        sig = torch.empty(NP, self.GH, self.GW, dtype=torch.float32, device=dev)
        logits = torch.empty(NP * self.GH * self.GW, 64, dtype=torch.float32, device=dev) if want_logits else None
        amax = torch.empty(NP, self.GH, self.GW, dtype=torch.int32, device=dev) if want_argmax else None
        ws = self._workspace(self.lib.kfn_workspace_bytes(self.h, 0, NP))
        _lib.check(self.lib.kfn_oflownet_forward(self.h, _ptr(f_prev), _ptr(f_cur), NP, _ptr(flow), _ptr(sig), _ptr(logits),
                                                 _ptr(amax), _ptr(ws), ws.numel(), self._stream()))
        return flow, sig, logits, amax

    # ---- one packed layer by name (per-layer tests / roofline) ----
    def conv_layer(self, key, x, out_hw, out_channels, out_coff=0, out=None, out_fp32=False, epi=0):
        N, IH, IW, ics = x.shape
        if key.endswith("#u8"):      # BF16-mode first layer straight from the u8 image (fused im2col producer)
            assert x.is_contiguous() and x.dtype == torch.uint8 and ics == 3
        else:
            assert x.is_contiguous() and x.dtype == (torch.float32 if self.mode == "fp32" else torch.bfloat16)
        odt = torch.float32 if (self.mode == "fp32" or out_fp32) else torch.bfloat16
        if out is None:
            out = torch.zeros(N, out_hw[0], out_hw[1], out_channels, dtype=odt, device=x.device)
        _lib.check(self.lib.kfn_conv_layer(self.h, key.encode(), _ptr(x), N, IH, IW, ics, _ptr(out), out.shape[1], out.shape[2],
                                           out.shape[3], out_coff, 1 if odt == torch.float32 else 0, epi, self._stream()))
        return out

    # ---- serial warp + Kalman scan over [S,T] ----
    def kalman_scan(self, measure, flow, sigma_w, transform, state_in=None, first_index=0, reinit_period=500,
                    nis_gate=False, diagnostics=False, want_state=False):
        S, T, GH, GW, _ = measure.shape
        dev = measure.device
        for t in (measure, flow, sigma_w):
            assert t.is_contiguous() and t.dtype == torch.float32
        if state_in is None and not (reinit_period > 0 and first_index % reinit_period == 0):
            raise ValueError("kalman_scan: state_in is required unless the first frame re-initialises the filter "
                             "(first_index % reinit_period == 0)")
        Tm = np.ascontiguousarray(np.asarray(transform, dtype=np.float32).reshape(16))
        out = torch.empty(S, T, GH, GW, 4, dtype=torch.float32, device=dev)
        temp = kf = nis = taps = state_out = None
        if diagnostics:
            temp = torch.empty_like(out); kf = torch.empty_like(out)
            nis = torch.empty(S, T, GH, GW, 3, dtype=torch.float32, device=dev)
            taps = torch.empty(S, T, GH, GW, 2, dtype=torch.int32, device=dev)
        if want_state:
            state_out = torch.empty(S, GH, GW, 4, dtype=torch.float32, device=dev)
        _lib.check(self.lib.kfn_kalman_scan(_ptr(measure), _ptr(flow), _ptr(sigma_w), _ptr(state_in), _ptr(state_out),
                                            Tm.ctypes.data_as(C.c_void_p), S, T, GH, GW, int(first_index), int(reinit_period),
                                            1 if nis_gate else 0, _ptr(out), _ptr(temp), _ptr(kf), _ptr(nis), _ptr(taps),
                                            self._stream()))
        res = dict(out_map=out, temp=temp, kf=kf, nis=nis, taps=taps, state=state_out)
        return res

    # ---- pointwise Kalman update / NIS on aligned maps (BuildKFCoord / GetNIS) ----
    def kalman_update(self, prior, measure, want_nis=False):
        assert prior.shape == measure.shape and prior.shape[-1] == 4 and prior.is_contiguous() and measure.is_contiguous()
        npix = prior.numel() // 4
        post = torch.empty_like(prior)
        nis = torch.empty(*prior.shape[:-1], 3, dtype=torch.float32, device=prior.device) if want_nis else None
        _lib.check(self.lib.kfn_kalman_update(_ptr(prior), _ptr(measure), npix, _ptr(post), _ptr(nis), self._stream()))
        return post, nis

    # ---- training (kfn_train_*; reference KFNet/train.py `run` + `solver`): FP32-mode engines only ----
    SCALAR_NAMES = ("loss", "measure_loss", "measure_accuracy", "temporal_loss", "temporal_accuracy", "kf_loss", "kf_accuracy",
                    "reg_loss", "lr")

    def train_begin(self, batch=4, train_scoordnet=True, train_oflownet=True):
        if self.mode != "fp32":
            raise ValueError("training runs on an fp32-mode Engine (the reference's arithmetic)")
        _lib.check(self.lib.kfn_train_begin(self.h, int(batch), 1 if train_scoordnet else 0, 1 if train_oflownet else 0))
        self._train_batch = int(batch)

    def train_step(self, images, gt8, mask8, base_lr=1e-4, weight_decay=1e-4, decay_steps=80000, gamma=0.5, apply_update=True):
        """images: cuda uint8 [B,H,W,3]; gt8: cuda fp32 [B,GH,GW,3] (label grid, network frame); mask8: cuda fp32 [B,GH,GW,1] in {0,1}.
        Returns the dict of scalars the reference's training loop prints (train.py:370-377) + reg_loss + lr."""
        B = getattr(self, "_train_batch", None)
        if B is None:
            raise RuntimeError("call train_begin first")
        self._chk_img(images)
        assert images.shape[0] == B and tuple(gt8.shape) == (B, self.GH, self.GW, 3) and tuple(mask8.shape) == (B, self.GH, self.GW, 1)
        for t in (gt8, mask8):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        scal = (C.c_float * 9)()
        _lib.check(self.lib.kfn_train_step(self.h, _ptr(images), _ptr(gt8), _ptr(mask8), float(base_lr), float(weight_decay),
                                           int(decay_steps), float(gamma), 1 if apply_update else 0, scal, self._stream()))
        self._graphs = {}          # parameters changed under any captured inference graph once committed
        return dict(zip(self.SCALAR_NAMES, [float(v) for v in scal]))

    def train_get(self, name, what="param"):
        """One variable (`param`), its last gradient (`grad`, without the regulariser term) or its Adam moments (`m`, `v`)."""
        shape = weight_shapes()[name]
        out = np.empty(shape, dtype=np.float32)
        _lib.check(self.lib.kfn_train_get(self.h, name.encode(), {"param": 0, "grad": 1, "m": 2, "v": 3}[what],
                                          out.ctypes.data_as(C.c_void_p), out.size))
        return out

    def train_set_step(self, step):
        """global_step.assign(step) (train.py:361-362,375; --reset_step)."""
        _lib.check(self.lib.kfn_train_set_step(self.h, int(step)))

    def train_commit(self):
        """Trained parameters -> the weight set the inference entry points use."""
        _lib.check(self.lib.kfn_train_commit(self.h))
        self._graphs = {}

    # ---- whole batched eval path on device-resident sequences ----
    def eval_sequences(self, images, transform, reinit_period=500, nis_gate=False, frames_per_batch=8, pairs_per_batch=4,
                       out=None):
        assert images.dim() == 5
        self._chk_img(images)
        S, T = images.shape[:2]
        Tm = np.ascontiguousarray(np.asarray(transform, dtype=np.float32).reshape(16))
        if out is None:
            out = torch.empty(S, T, self.GH, self.GW, 4, dtype=torch.float32, device=images.device)
        need = (self.lib.kfn_workspace_bytes(self.h, frames_per_batch, pairs_per_batch)
                + self.lib.kfn_sequence_extra_bytes(self.h, S, T))
        ws = self._workspace(need)
        _lib.check(self.lib.kfn_eval_sequences(self.h, _ptr(images), S, T, Tm.ctypes.data_as(C.c_void_p), int(reinit_period),
                                               1 if nis_gate else 0, int(frames_per_batch), int(pairs_per_batch), _ptr(out),
                                               _ptr(ws), ws.numel(), self._stream()))
        return out
