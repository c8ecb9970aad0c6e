"""Per-launch roofline table from a committed ncu launch list (profiles/<tag>_launches.csv, produced by
`ncu --metrics gpu__time_duration.sum ... bench.py --steps 1 --warmup 0 --frames 16 --graph 0`): one 16-frame batch,
one 15-pair batch and the scan.  Algorithmic FLOPs = 2 x MACs of the layer as the reference defines it (padded taps
counted); algorithmic bytes = activations in + out (bf16 unless noted) + weights once.  ncu serialises
launches and measures each cold, so times are upper bounds of the in-step times; peaks are MEASURED_PEAKS.json
(bf16 burst, HBM copy).  CPU only.
  python tools/per_layer_roofline.py r01i > profiles/r01i_per_layer_roofline.md"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01i"
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
TF, HBM = pk["bf16_tflops"], pk["hbm_gbs"]
B, NP, H, W = 16, 15, 480, 640
R = NP * 60 * 80                                   # U-Net windows per pair batch


def conv(name, px, k, cin, cout, in_px=None, in_b=2, out_b=2, out_c=None):
    macs = px * k * k * cin * cout
    in_px = px if in_px is None else in_px
    byt = in_px * cin * in_b + px * (cout if out_c is None else out_c) * out_b + k * k * cin * cout * 2
    return (name, 2 * macs, byt)


full, half, quar, eig = B * H * W, B * H * W // 4, B * H * W // 16, B * H * W // 64
layers = [
    ("first layer: conv1a + feat1 from u8 (N = 80)", 2 * full * 27 * 80, full * 3 + full * 80 * 2 + 80 * 32 * 2),
    conv("ScoreNet/conv1b", full, 3, 64, 64), conv("ScoreNet/conv2a s2", half, 3, 64, 256, in_px=full),
    conv("ScoreNet/conv2b", half, 3, 256, 256), conv("ScoreNet/conv3a s2", quar, 3, 256, 512, in_px=half),
    conv("ScoreNet/conv3b", quar, 3, 512, 512), conv("ScoreNet/conv4a s2", eig, 3, 512, 1024, in_px=quar),
    conv("ScoreNet/conv4b", eig, 3, 1024, 1024), conv("ScoreNet/conv5", eig, 3, 1024, 512),
    conv("ScoreNet/conv6", eig, 3, 512, 256), conv("ScoreNet/conv7 1x1", eig, 1, 256, 128),
    conv("ScoreNet/prediction 1x1 (+exp)", eig, 1, 128, 4, out_b=4),
    conv("Temporal/feat2 s2", half, 3, 16, 32, in_px=full), conv("Temporal/feat3", half, 3, 32, 32),
    conv("Temporal/feat4 s2", quar, 3, 32, 64, in_px=half), conv("Temporal/feat5", quar, 3, 64, 64),
    conv("Temporal/feat6 s2", eig, 3, 64, 128, in_px=quar), conv("Temporal/feat7 (+l2norm)", eig, 3, 128, 32, out_b=4),
    ("difference volume", 0, NP * (2 * 4800 * 32 * 4 + 4800 * 64 * 32 * 2)),
    conv("U-Net/conv0", R * 64, 3, 32, 32), conv("U-Net/conv1a s2", R * 16, 3, 32, 32, in_px=R * 64),
    conv("U-Net/conv1b", R * 16, 3, 32, 32), conv("U-Net/conv2a s2", R * 4, 3, 32, 64, in_px=R * 16),
    conv("U-Net/conv2b", R * 4, 3, 64, 64), conv("U-Net/conv3a s2", R, 3, 64, 128, in_px=R * 4),
    conv("U-Net/conv3b", R, 3, 128, 128),
    ("U-Net/upconv2 (4 phases, N = 256)", 2 * R * 9 * 128 * 64, R * 128 * 2 + R * 4 * 64 * 2 + 9 * 128 * 64 * 2),
    conv("U-Net/conv4", R * 4, 3, 128, 64),
    ("U-Net/upconv1 (4 phases, N = 128)", 2 * R * 4 * 9 * 64 * 32, R * 4 * 64 * 2 + R * 16 * 32 * 2 + 9 * 64 * 32 * 2),
    conv("U-Net/conv5", R * 16, 3, 64, 32),
    ("U-Net/upconv0 (4 phases, N = 64)", 2 * R * 16 * 9 * 32 * 16, R * 16 * 32 * 2 + R * 64 * 16 * 2 + 9 * 32 * 16 * 2),
    conv("U-Net/conv6", R * 64, 3, 48, 16), conv("U-Net/prediction", R * 64, 3, 16, 1, out_b=4),
    conv("FC fc1", R, 1, 128, 64), conv("FC fc2", R, 1, 64, 32), conv("FC uncertainty", R, 1, 32, 1, out_b=4),
    ("softmax + flow", 0, R * (64 * 4 + 4 + 8 + 4)),
    ("Kalman scan (16 frames, thread-block cluster)", 0, 16 * 364800),
]
rows = list(csv.DictReader(open(os.path.join(ROOT, "profiles", tag + "_launches.csv"))))[:len(layers)]
print("| # | layer (one launch) | kernel | µs | GFLOP | TFLOP/s | % bf16 burst peak | MB | GB/s | % HBM peak |")
print("|---|---|---|---|---|---|---|---|---|---|")
tot_us = tot_fl = 0.0
for i, ((name, fl, by), r) in enumerate(zip(layers, rows)):
    us = int(r["duration_ns"]) / 1e3
    tot_us += us; tot_fl += fl
    tfs = fl / us / 1e6
    gbs = by / us / 1e3
    print("| %d | %s | %s | %.1f | %.1f | %.0f | %.0f | %.1f | %.0f | %.0f |" % (
        i, name, r["kernel"].replace("unnamed>::", "").replace("conv_tc_kernel", "tc").replace("void ", ""), us, fl / 1e9, tfs, 100 * tfs / TF, by / 1e6, gbs, 100 * gbs / HBM))
print()
print("Batch total: %.0f µs under ncu (serialised, cold), %.1f GFLOP -> %.0f TFLOP/s overall = %.0f %% of the burst bf16 peak." % (
    tot_us, tot_fl / 1e9, tot_fl / tot_us / 1e6, 100 * tot_fl / tot_us / 1e6 / TF))
