"""BASELINE.json configs[0..2] on one B200: numbers + parity next to each other -> gpurun_out/configs_report.json
  configs[0]: single 480x640 frame, SCoordNet-only on CPU (oracle, torch-CPU fp32)
  configs[1]: SCoordNet conv stack as tcgen05/TMA kernel, batch 8, fp32-grade (bf16x3) vs the fp64 oracle; bf16 and fp32-SIMT beside it
  configs[2]: OFlowNet difference volume + flow-warp / Kalman kernels on one pair: achieved HBM GB/s"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

g.build()
from kfnet_b200 import synth  # noqa: E402
from kfnet_b200.engine import Engine  # noqa: E402
from oracle import kfnet_oracle as O  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
w = synth.make_weights(0)
imgs = synth.make_sequence(8, seed=1)
dev = torch.from_numpy(imgs).cuda()
rep = {}


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    s = b.pow(2).mean().sqrt()
    return float(((a - b).abs() / torch.maximum(b.abs(), s)).max())


def gpu_ms(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# configs[0]
t0 = time.perf_counter(); c32, u32 = O.scoordnet(imgs[:1].astype(np.float32), w, torch.float32); dt = time.perf_counter() - t0
c64, u64 = O.scoordnet(imgs[:1].astype(np.float32), w, torch.float64)
rep["config0_cpu_scoordnet_1frame"] = {"seconds": dt, "threads": torch.get_num_threads(), "fp32_vs_fp64_err": rel(c32, c64)}
# configs[1]
rows = {}
for mode in ("bf16x3", "bf16", "fp32"):
    eng = Engine(w, mode=mode)
    ms = gpu_ms(lambda: eng.scoordnet(dev), reps=3 if mode == "fp32" else 10)
    m = eng.scoordnet(dev)
    rows[mode] = {"ms_per_batch8": ms, "frames_per_s": 8e3 / ms, "TFLOPs_algorithmic": 465.69 * 8 / ms,
                  "coord_err_vs_fp64_oracle": rel(m[:1, ..., :3], c64), "log_sigma_err": rel(m[:1, ..., 3:4].log(), u64.log())}
    eng.close()
rep["config1_scoordnet_batch8"] = rows
# configs[2]
eng = Engine(w, mode="bf16")
feat = eng.oflow_feat(dev[:2])
fp, fc = feat[:1].contiguous(), feat[1:].contiguous()
out16 = eng.diff_volume(fp, fc, bf16=True); out32 = eng.diff_volume(fp, fc, bf16=False)
ms16 = gpu_ms(lambda: eng.diff_volume(fp, fc, bf16=True, out=out16), 50)
ms32 = gpu_ms(lambda: eng.diff_volume(fp, fc, bf16=False, out=out32), 50)
b16, b32 = 1228800 + 19660800, 1228800 + 39321600
S, T = 8, 256
gsc = torch.Generator().manual_seed(1)
meas = torch.randn(S, T, 60, 80, 4, generator=gsc).cuda(); meas[..., 3] = 0.1
flow = (torch.rand(S, T, 60, 80, 2, generator=gsc) * 2 - 1).cuda(); sw = torch.full((S, T, 60, 80), 0.01).cuda()
Tm = np.eye(4, dtype=np.float32)
ms_scan = gpu_ms(lambda: eng.kalman_scan(meas, flow, sw, Tm, reinit_period=500), 5)
scan_bytes = S * T * 4800 * (16 + 8 + 4 + 16)          # measure + flow + sigma_w read, out map written
# throughput mode of the scan: one sequence per SM (148 CTAs), the regime in which its HBM rate is meaningful
S2, T2 = 148, 32
meas2 = torch.randn(S2, T2, 60, 80, 4, generator=gsc).cuda(); meas2[..., 3] = 0.1
flow2 = (torch.rand(S2, T2, 60, 80, 2, generator=gsc) * 2 - 1).cuda(); sw2 = torch.full((S2, T2, 60, 80), 0.01).cuda()
ms_scan2 = gpu_ms(lambda: eng.kalman_scan(meas2, flow2, sw2, Tm, reinit_period=500), 3)
scan_bytes2 = S2 * T2 * 4800 * (16 + 8 + 4 + 16)
rep["config2_pair_kernels"] = {
    "diff_volume_bf16": {"us": ms16 * 1e3, "GBps": b16 / ms16 / 1e6, "frac_hbm": b16 / ms16 / 1e6 / peaks["hbm_gbs"], "bytes": b16},
    "diff_volume_fp32": {"us": ms32 * 1e3, "GBps": b32 / ms32 / 1e6, "frac_hbm": b32 / ms32 / 1e6 / peaks["hbm_gbs"], "bytes": b32},
    "kalman_scan_8seq_x_256frames": {"us_per_frame": ms_scan * 1e3 / T, "GBps": scan_bytes / ms_scan / 1e6,
                                     "frac_hbm": scan_bytes / ms_scan / 1e6 / peaks["hbm_gbs"], "bytes_per_frame_per_seq": 4800 * 44,
                                     "note": "serial over frames by construction (state in smem); 8 CTAs = 8 SMs of 148"},
    "kalman_scan_148seq_x_32frames": {"us_per_frame_step": ms_scan2 * 1e3 / T2, "GBps": scan_bytes2 / ms_scan2 / 1e6,
                                      "frac_hbm": scan_bytes2 / ms_scan2 / 1e6 / peaks["hbm_gbs"],
                                      "note": "one sequence per SM: all 148 SMs busy"},
}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rep, open("gpurun_out/configs_report.json", "w"), indent=1)
print(json.dumps(rep, indent=1))
