"""Experiment: per-frame latency of the warp + Kalman scan for one / eight sequences (KFN_SCAN_CLUSTER_SIZE from the environment).  GPU only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from kfnet_b200.engine import Engine
from kfnet_b200.train import initial_weights

eng = Engine(initial_weights(0), mode="bf16", image_size=(480, 640))
g = torch.Generator().manual_seed(0)
for S in (1, 8):
    T = 256
    measure = torch.randn(S, T, 60, 80, 4, generator=g).cuda().contiguous()
    measure[..., 3] = measure[..., 3].abs() + 0.1
    flow = (torch.rand(S, T, 60, 80, 2, generator=g) * 6 - 3).cuda().contiguous()
    sig = (torch.rand(S, T, 60, 80, generator=g) * 0.1 + 0.01).cuda().contiguous()
    tr = np.eye(4, dtype=np.float32)
    fn = lambda: eng.kalman_scan(measure, flow, sig, tr)
    for _ in range(3):
        r = fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        r = fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print("cluster size %s, S=%d: %.3f ms per %d-frame scan = %.2f us per frame; checksum %.6f" % (
        os.environ.get("KFN_SCAN_CLUSTER_SIZE", "default"), S, ms, T, ms * 1000 / T, float(r["out_map"].double().sum())))
