#!/bin/bash
# compute-sanitizer over the code paths added in round 2: training kernels, fp16x2 / halo-x3 / fused-limb stages, streaming eval.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo "build failed"; exit 1; }
S=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $S --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_training.py -q -m gpu -x -k "two_adam or frozen" > $O/san_memcheck_training.log 2>&1; echo "memcheck training exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $O/san_memcheck_training.log | head -n 6
timeout 900 $S --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_networks.py -q -m gpu -x -k "eval_sequences_small or scoordnet_and_tower_small" > $O/san_memcheck_networks.log 2>&1; echo "memcheck networks(small: all 4 modes incl. fp16x2 / bf16x3 halo + fused-limb stages) exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $O/san_memcheck_networks.log | head -n 6
timeout 900 $S --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_eval_dropin.py -q -m gpu -x -k "eval_cli" > $O/san_memcheck_dropin.log 2>&1; echo "memcheck drop-in eval (kfn_eval_stream_host) exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $O/san_memcheck_dropin.log | head -n 6
timeout 900 $S --tool racecheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_training.py -q -m gpu -x -k "frozen" > $O/san_racecheck_training.log 2>&1; echo "racecheck training exit $?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" $O/san_racecheck_training.log | head -n 6
# ncu: what limits the N = 256 / 512 layers (conv2a .. conv3b = conv_tc_kernel launches 0..3 of a batch), batch 16
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'conv_tc_kernel$' -c 4 -f -o $O/prof_big_layers \
    python bench.py --steps 1 --warmup 0 --frames 16 --frames_per_batch 16 --pairs_per_batch 16 --cpu_frames 0 --graph 0 --parity_frames 0 --extras 0 --ncu 0 --dropin_frames 0 > $O/ncu_big.log 2>&1; echo "ncu big layers exit $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_r02i.csv \
    python bench.py --steps 1 --warmup 0 --frames 16 --frames_per_batch 16 --pairs_per_batch 16 --cpu_frames 0 --graph 0 --parity_frames 0 --extras 0 --ncu 0 --dropin_frames 0 > $O/ncu_list_r02i.log 2>&1; echo "ncu list exit $?"
