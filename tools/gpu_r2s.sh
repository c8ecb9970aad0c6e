#!/bin/bash
# Round 2, call S: ncu --set full of the small-N layers (tower feat2..7, U-Net, FC) of one 16-frame batch: where do the
# layers that sit at 4-30 % of the tensor peak spend their time?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo "build failed"; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_tc_kernel(_x2(_lean)?)?$' -c 26 -f -o $O/prof_small \
    python bench.py --steps 1 --warmup 0 --frames 16 --frames_per_batch 16 --pairs_per_batch 16 --cpu_frames 0 --graph 0 --parity_frames 0 --extras 0 --ncu 0 --dropin_frames 0 > $O/ncu_small.log 2>&1; echo "ncu small exit $?"
tail -3 $O/ncu_small.log
