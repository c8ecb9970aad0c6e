"""Experiment: does the interleaved concat layout (upconv0's 16 channels next to conv0's 32 in 96-byte pixels) cost the
U-Net's 8x8-level layers time?  Times upconv0 / conv0 writing into the concat buffer vs into compact tensors.  GPU only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kfnet_b200.engine import Engine
from kfnet_b200.train import initial_weights

R = 15 * 4800
eng = Engine(initial_weights(0), mode="bf16", image_size=(480, 640))
g = torch.Generator().manual_seed(0)
x5 = torch.randn(R, 4, 4, 32, generator=g).cuda().to(torch.bfloat16).contiguous()
xd = torch.randn(R, 8, 8, 32, generator=g).cuda().to(torch.bfloat16).contiguous()


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1000.0


cat = torch.zeros(R, 8, 8, 48, dtype=torch.bfloat16, device="cuda")
up_c = torch.zeros(R, 8, 8, 16, dtype=torch.bfloat16, device="cuda")
c0_c = torch.zeros(R, 8, 8, 32, dtype=torch.bfloat16, device="cuda")
c6 = torch.zeros(R, 8, 8, 16, dtype=torch.bfloat16, device="cuda")
print("upconv0 -> concat buffer (48 ch, coff 0): %.1f us" % timeit(lambda: eng.conv_layer("Temporal/upconv0#all", x5, (8, 8), 48, out_coff=0, out=cat)))
print("upconv0 -> compact 16 ch               : %.1f us" % timeit(lambda: eng.conv_layer("Temporal/upconv0#all", x5, (8, 8), 16, out_coff=0, out=up_c)))
print("conv0   -> concat buffer (48 ch, coff 16): %.1f us" % timeit(lambda: eng.conv_layer("Temporal/conv0", xd, (8, 8), 48, out_coff=16, out=cat)))
print("conv0   -> compact 32 ch                : %.1f us" % timeit(lambda: eng.conv_layer("Temporal/conv0", xd, (8, 8), 32, out_coff=0, out=c0_c)))
print("conv6   <- concat buffer                : %.1f us" % timeit(lambda: eng.conv_layer("Temporal/conv6", cat, (8, 8), 16, out_coff=0, out=c6)))
