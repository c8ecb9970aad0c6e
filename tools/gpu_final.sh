#!/bin/bash
# Round-end evidence run on one B200: tests, smoke, bench (N=1), reference arm, ncu launch list + full captures.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; tail -n 3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json | cut -c1-300
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference exit $?"; cat gpurun_out/bench_reference.json | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 0 --frames 16 --cpu_frames 0 --graph 0 > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
# full captures: conv4b (launch 8 of conv_tc), diff volume, softmax-flow, kalman scan
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 7 -c 2 -f -o gpurun_out/prof_conv_tc \
    python bench.py --steps 1 --warmup 0 --frames 16 --cpu_frames 0 --graph 0 > gpurun_out/ncu_full1.log 2>&1; echo "ncu conv exit $?"
ncu --set full --clock-control none --import-source on -k regex:"diff_volume|kalman_scan|softmax_flow" -c 3 -f -o gpurun_out/prof_pointwise \
    python bench.py --steps 1 --warmup 0 --frames 16 --cpu_frames 0 --graph 0 > gpurun_out/ncu_full2.log 2>&1; echo "ncu pointwise exit $?"
