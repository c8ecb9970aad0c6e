#!/bin/bash
# Round 2, call P: ncu --set full of the CTA-pair kernel (conv2a .. conv6 of one 16-frame batch) for the committed evidence.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo "build failed"; exit 1; }
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'conv_tc_kernel_pair' -c 8 -f -o $O/prof_pair \
    python bench.py --steps 1 --warmup 0 --frames 16 --frames_per_batch 16 --pairs_per_batch 16 --cpu_frames 0 --graph 0 --parity_frames 0 --extras 0 --ncu 0 --dropin_frames 0 > $O/ncu_pair.log 2>&1; echo "ncu pair exit $?"
