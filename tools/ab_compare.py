"""Bit-level A/B of kernel-variant switches (they are read once per process): run the BF16 OFlowNet + SCoordNet
forward on fixed synthetic inputs and save the outputs; run again with different KFN_TC_* settings; compare.
  KFN_TC_DIFF_FUSED=0 python tools/ab_compare.py gpurun_out/ab_a.npz
  KFN_TC_DIFF_FUSED=1 python tools/ab_compare.py gpurun_out/ab_b.npz gpurun_out/ab_a.npz     # prints max |a-b| per output
Variants that only move where an operand is formed (fused u8 first layer, fused difference volume, fused conv1a+conv1b)
must give identical bits."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from kfnet_b200 import synth  # noqa: E402
from kfnet_b200.engine import Engine  # noqa: E402


def main():
    out_path = sys.argv[1]
    eng = Engine(synth.make_weights(0), mode="bf16")
    imgs = torch.from_numpy(synth.make_sequence(5, seed=2)).cuda()
    meas = eng.scoordnet(imgs)
    feat = eng.oflow_feat(imgs)
    flow, sigw, logits, amax = eng.oflownet(feat[:-1].contiguous(), feat[1:].contiguous(), want_logits=True, want_argmax=True)
    torch.cuda.synchronize()
    res = dict(measure=meas.cpu().numpy(), feat=feat.cpu().numpy(), flow=flow.cpu().numpy(), sigma_w=sigw.cpu().numpy(),
               logits=logits.cpu().numpy(), argmax=amax.cpu().numpy())
    np.savez(out_path, **res)
    print("saved", out_path, {k: v.shape for k, v in res.items()})
    if len(sys.argv) > 2:
        ref = np.load(sys.argv[2])
        worst = 0.0
        for k, v in res.items():
            d = float(np.abs(v.astype(np.float64) - ref[k].astype(np.float64)).max())
            worst = max(worst, d)
            print("%-8s max |a-b| = %.3e  %s" % (k, d, "identical" if d == 0 else ""))
        sys.exit(0 if worst == 0 else 1)


if __name__ == "__main__":
    main()
