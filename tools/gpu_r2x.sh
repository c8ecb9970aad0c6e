#!/bin/bash
# Round 2, call X: training tests on the final source (train_begin error path), ncu --set full of the U-Net launches of the final build.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo "build failed"; exit 1; }
timeout 600 python -m pytest tests/test_gpu_training.py -q -m gpu > $O/pytest_x.log 2>&1; echo "pytest (training) exit $?"; grep -E "passed|failed|^E  |FAILED" $O/pytest_x.log | cut -c1-300 | tail -n 6
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 $O/smoke.log | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'conv_tc_kernel(_x2(_lean(_ph)?)?)?$' -s 7 -c 14 -f -o $O/prof_unet_final \
    python bench.py --steps 1 --warmup 0 --frames 16 --frames_per_batch 16 --pairs_per_batch 16 --cpu_frames 0 --graph 0 --parity_frames 0 --extras 0 --ncu 0 --dropin_frames 0 > $O/ncu_unet.log 2>&1; echo "ncu exit $?"
