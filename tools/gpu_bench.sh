#!/bin/bash
# One gpurun call: build, smoke, bench (N=1), ncu launch list of a short bench, optional full ncu capture.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 4 gpurun_out/smoke.log
python bench.py --steps ${STEPS:-3} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
if [ -n "$NCU_LIST" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 0 --frames 16 --cpu_frames 0 > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
fi
if [ -n "$NCU_FULL" ]; then
  ncu --set full --clock-control none --import-source on -k regex:conv_tc -s ${NCU_SKIP:-3} -c ${NCU_COUNT:-4} -f -o gpurun_out/prof_conv_tc \
      python bench.py --steps 1 --warmup 0 --frames 8 --cpu_frames 0 > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
fi
