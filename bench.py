"""bench.py -- KFNet hot path on B200: frames/s at 480x640 (60x80 label grid).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode bf16|fp32] [--frames T]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

Workload (BASELINE.json configs[3] per GPU, configs[4] over N GPUs): full recurrent KFNet eval -- SCoordNet +
OFlowNet feature tower + difference volume + U-Net/FC + soft-argmax flow + warp + Kalman scan -- of one
synthetic 256-frame 480x640 sequence per rank, random weights (kfnet_b200/synth.py), bf16 compute / fp32
accumulate.  One step = one pass of the whole path over the rank's sequence.  N ranks = N independent
sequences (weak scaling) + one NCCL all_gather of the per-rank output maps inside the timed step.

  value : frames/s with the uint8 images already resident in HBM (CUDA events, max over ranks)
  e2e   : same metric through ONE C-ABI call per step with HOST (pinned) buffers (kfn_eval_sequences_host): H2D of the
          images (streamed batch by batch inside the call), the whole path and the D2H of the [T,60,80,4] output maps
          inside the timed region, every step; `e2e_pipelined` additionally overlaps the H2D of step i+1 with step i
  --impl reference : the reference's own per-frame eval loop (batch-of-2 SCoordNet+tower recomputation
          included, KFNet/eval.py:41,77-104) restated op-for-op on torch-CPU (oracle/), timed on the host cores.
          TensorFlow 1.13 cannot be installed offline, so this is kind="port".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic work per 480x640 frame (2 flops per MAC, padded taps counted)
GFLOP_SCOORDNET = 465.69
GFLOP_TOWER = 5.574
GFLOP_OFLOWNET = 24.96      # U-Net (conv6 has 48 input channels) + FC head, per frame pair


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        top = sorted(sm)[len(sm) // 2:]          # upper half = samples under load
        return {"sm_mhz": statistics.median(top), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def reference_loop_fps(frames, threads=None, seed=1):
    """Reference-faithful CPU iteration (oracle, faithful_redundancy=True) on `frames` synthetic frames."""
    import numpy as np
    import torch
    from kfnet_b200 import synth
    from oracle import kfnet_oracle as O
    if threads:
        torch.set_num_threads(threads)
    w = synth.make_weights(0)
    imgs = synth.make_sequence(frames, seed=seed).astype(np.float32)
    Tm = np.linalg.inv(synth.make_transform(2)).astype(np.float32)
    t0 = time.perf_counter()
    with torch.no_grad():
        O.eval_sequence(imgs, w, Tm, seq_len=500, dtype=torch.float32, faithful_redundancy=True)
    dt = time.perf_counter() - t0
    return frames / dt, dt, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count()
    per_step = 2
    for _ in range(args.warmup):
        reference_loop_fps(per_step)
    t0 = time.perf_counter()
    for k in range(args.steps):
        reference_loop_fps(per_step, seed=1 + k)
    dt = time.perf_counter() - t0
    fps = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": "frames/sec at 480x640 (60x80 label grid)", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "full KFNet recurrent eval, 480x640 synthetic frames, reference per-frame loop "
                               "(batch-of-2 recomputation, eval.py:41,77-104); each step = %d frames" % per_step},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "%d steps x %d frames of the reference eval iteration, torch-CPU fp32 (oneDNN), "
                                   "host has %d logical cores; TF1 not installable offline" % (args.steps, per_step, cores)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    from kfnet_b200 import synth
    from kfnet_b200.engine import Engine
    from kfnet_b200.shard import gather_outputs

    T, H, W = args.frames, 480, 640
    weights = synth.make_weights(0)
    Tm = np.linalg.inv(synth.make_transform(2)).astype(np.float32)
    eng = Engine(weights, mode=args.mode, image_size=(H, W), device=local)
    # synthetic sequence of this rank (seed = rank + 1), generated in chunks to bound host memory
    host_imgs = torch.empty((1, T, H, W, 3), dtype=torch.uint8).pin_memory()
    base = synth.make_sequence(min(T, 32), seed=rank + 1)
    for t0 in range(0, T, base.shape[0]):
        n = min(base.shape[0], T - t0)
        host_imgs[0, t0:t0 + n] = torch.from_numpy(np.roll(base[:n], shift=t0 // base.shape[0], axis=2))
    dev_imgs = host_imgs.cuda(non_blocking=False)
    out = torch.empty(1, T, eng.GH, eng.GW, 4, dtype=torch.float32, device="cuda")
    host_out = torch.empty((1, T, eng.GH, eng.GW, 4), dtype=torch.float32).pin_memory()
    fpb, ppb = args.frames_per_batch, args.pairs_per_batch

    def run_eval(imgs):
        if args.graph:
            eng.eval_sequences_graphed(imgs, Tm, out, reinit_period=500, frames_per_batch=fpb, pairs_per_batch=ppb)
        else:
            eng.eval_sequences(imgs, Tm, reinit_period=500, frames_per_batch=fpb, pairs_per_batch=ppb, out=out)

    def step_resident():
        run_eval(dev_imgs)
        if world > 1:
            gather_outputs(out, world)          # the path's single collective: all_gather of the per-rank output maps

    # end-to-end path: pinned host images -> device -> eval -> pinned host maps, every step.  Input copies are
    # double-buffered on a copy stream (the H2D of step i+1 overlaps the compute of step i, as a caller streaming
    # sequences would do); each step still performs exactly one H2D of its images and one D2H of its maps.
    stages = [torch.empty_like(dev_imgs), torch.empty_like(dev_imgs)]
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    free = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_i = [0]

    def issue_h2d(k):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free[k])
            stages[k].copy_(host_imgs, non_blocking=True)
            ready[k].record(copy_stream)

    for k in range(2):
        free[k].record()
    issue_h2d(0)

    def step_e2e():
        k = e2e_i[0] & 1
        e2e_i[0] += 1
        issue_h2d(k ^ 1)                                  # next step's input, overlapped with this step's compute
        torch.cuda.current_stream().wait_event(ready[k])
        run_eval(stages[k])
        free[k].record()
        host_out.copy_(out, non_blocking=True)
        if world > 1:
            gather_outputs(out, world)

    def step_e2e_cabi():
        # ONE C-ABI call with HOST buffers (kfn_eval_sequences_host): H2D of the frames batch by batch on the library's
        # copy stream, the whole path, D2H of the maps, stream synchronised before it returns
        eng.eval_sequences_host(host_imgs, Tm, host_out, reinit_period=500, frames_per_batch=fpb, pairs_per_batch=ppb)
        if world > 1:
            gather_outputs(eng._host_stage[1], world)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # kernels per step: counted on one eager step (graph replays launch the same kernels without passing the counter)
    eng.launch_count(reset=True)
    eng.eval_sequences(dev_imgs, Tm, reinit_period=500, frames_per_batch=fpb, pairs_per_batch=ppb, out=out)
    launches = eng.launch_count()
    ms = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e_pipe = timed(step_e2e, args.steps, max(4, args.warmup))      # >= 4: both staging buffers get their graph captured in warm-up
    ms_e2e = timed(step_e2e_cabi, args.steps, args.warmup)

    # per-stage device times on this rank (events on the launching stream), for the roofline of the dominant kernel
    def stage_ms(fn, reps=3):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    B = fpb
    t_sc = stage_ms(lambda: eng.scoordnet(dev_imgs[0, :B]))
    eng.launch_count(reset=True); eng.scoordnet(dev_imgs[0, :B]); n_sc = eng.launch_count()
    t_tw = stage_ms(lambda: eng.oflow_feat(dev_imgs[0, :B]))
    feat = eng.oflow_feat(dev_imgs[0, :ppb + 1])
    t_of = stage_ms(lambda: eng.oflownet(feat[:-1].contiguous(), feat[1:].contiguous()))
    fp_, fc_ = feat[:-1].contiguous(), feat[1:].contiguous()
    dv_out = eng.diff_volume(fp_, fc_, bf16=(args.mode != "fp32"))
    t_dv = stage_ms(lambda: eng.diff_volume(fp_, fc_, bf16=(args.mode != "fp32"), out=dv_out), reps=10)
    pk = peaks()
    tf_ach = GFLOP_SCOORDNET * B / t_sc            # GFLOP / ms = TFLOP/s
    # the dominant single launch: conv4b (3x3, 1024->1024 on the 60x80 grid), timed per launch through kfn_conv_layer
    roof_launch = None
    if args.mode == "bf16":
        x4 = (torch.randn(B, eng.GH, eng.GW, 1024, device="cuda") * 0.2).to(torch.bfloat16)
        y4 = torch.empty(B, eng.GH, eng.GW, 1024, device="cuda", dtype=torch.bfloat16)
        t_c4 = stage_ms(lambda: eng.conv_layer("ScoreNet/conv4b", x4, (eng.GH, eng.GW), 1024, out=y4), reps=20)
        gflop_c4 = 2.0 * B * eng.GH * eng.GW * 9 * 1024 * 1024 / 1e9
        traffic = None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum of the same launch from the committed `ncu --set full` capture
            import csv
            import glob
            f = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_prof_conv_tc_ncu_summary.csv")))[-1]
            rows = list(csv.reader(open(f)))
            d = dict(zip(rows[0], rows[2]))
            traffic = {"bytes": (float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"])) * 1e6,
                       "algorithmic_bytes": B * eng.GH * eng.GW * 1024 * 2 * 2 + 9 * 1024 * 1024 * 2,
                       "tensor_pipe_active_pct": float(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]),
                       "source": os.path.basename(f) + " (conv4b at batch 16)"}
        except Exception:
            pass
        roof_launch = {"bound": "tensor", "kernel": "conv_tc_kernel, ScoreNet/conv4b, one launch, batch %d" % B,
                       "achieved": gflop_c4 / t_c4, "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": gflop_c4 / t_c4 / pk["tf_burst"],
                       "peak_source": pk["src"] + " bf16 burst (kernel timed alone)", "flops_per_launch": gflop_c4 * 1e9,
                       "ms_per_launch": t_c4, "traffic": traffic}
        del x4, y4
    dv_bytes = ppb * (2 * 60 * 80 * 32 * 4 + 4800 * 64 * 32 * (2 if args.mode == "bf16" else 4))
    frames_total = world * T * args.steps
    fps = frames_total / (ms / 1e3)
    fps_e2e = frames_total / (ms_e2e / 1e3)

    if rank == 0:
        cpu_fps, cpu_dt, cpu_threads = reference_loop_fps(args.cpu_frames) if args.cpu_frames > 0 else (None, 0, 0)
        line = {
            "metric": "frames/sec at 480x640 (60x80 label grid)", "value": fps, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": {"bf16": "bf16", "bf16x3": "bf16x3 (3 bf16 limbs, fp32 accumulate)", "fp32": "f32"}[args.mode], "data": "synthetic",
            "config": {"workload": "full KFNet recurrent eval (SCoordNet + OFlowNet tower/diff-volume/U-Net + warp + Kalman scan), "
                                   "one %d-frame synthetic 480x640 sequence per GPU (BASELINE configs[3]; configs[4] over N GPUs), "
                                   "random weights" % T,
                       "frames_per_batch": fpb, "pairs_per_batch": ppb, "compute": args.mode, "cuda_graph": bool(args.graph),
                       "parallelism": "sequence-sharded x%d, one all_gather of output maps per step" % world,
                       "l2": "inputs larger than L2 (%.0f MB of uint8 images per step)" % (T * H * W * 3 / 1e6)},
            "clocks": clocks,
            "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": int(T * H * W * 3),
                    "d2h_bytes_per_step": int(T * eng.GH * eng.GW * 16),
                    "api": "kfn_eval_sequences_host (one C-ABI call per step, pinned host buffers in/out, synchronous)"},
            "e2e_pipelined": {"value": frames_total / (ms_e2e_pipe / 1e3), "unit": "frames/s",
                              "api": "Engine.eval_sequences_graphed + double-buffered H2D across steps (CUDA graph replay)"},
            "gpu_launches": int(launches * args.steps),
            "roofline": {"bound": "tensor", "kernel": "conv_tc kernels over the SCoordNet stack (%d launches per batch of %d frames, first layer reads the u8 frames; ~60 %% of the step)" % (n_sc, B),
                         "achieved": tf_ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": tf_ach / pk["tf_sustained"],
                         "peak_source": pk["src"] + " bf16 sustained", "traffic": None},
            "roofline_dominant_launch": roof_launch,
            "stages_ms": {"scoordnet_per_batch": t_sc, "tower_per_batch": t_tw, "oflownet_per_pair_batch": t_of,
                          "diff_volume_per_pair_batch": t_dv,
                          "diff_volume_GBps": dv_bytes / t_dv / 1e6, "diff_volume_frac_hbm": dv_bytes / t_dv / 1e6 / pk["hbm"]},
            "cpu_baseline": ({"value": cpu_fps, "unit": "frames/s", "cores": cpu_threads, "kind": "port",
                              "sample": "%d frames of the reference eval iteration (batch-of-2 recomputation) on torch-CPU fp32, %.1f s; "
                                        "host has %d logical cores; TF1 not installable offline" % (args.cpu_frames, cpu_dt, os.cpu_count())}
                             if cpu_fps else None),
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def _capture_stdout():
    """Everything native libraries print to fd 1 (e.g. NCCL's version banner) goes to stderr; the one JSON line is
    written to the saved real stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        sys.stdout = sys.stderr


def emit(line: dict):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32", "bf16x3"])
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--frames_per_batch", type=int, default=16)
    ap.add_argument("--pairs_per_batch", type=int, default=16)
    ap.add_argument("--cpu_frames", type=int, default=6)
    ap.add_argument("--graph", type=int, default=1, help="replay each step from a CUDA graph (0 = eager launches)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
