#!/bin/bash
# Round 2, call B: new fp16x2 mode + API hardening -> full GPU suite, smoke, the rewritten bench (all blocks), reference arm,
# per-launch lists of the error-compensated modes.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo "build failed"; tail -5 $O/build.log; exit 1; }
timeout 600 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; tail -n 15 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -n 5 $O/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; tail -c 600 $O/bench.err; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench.json"))
    for k in ("value", "ms_per_step", "e2e", "roofline", "roofline_whole_step", "parity_mode", "configs4_on_one_gpu", "cpu_baseline", "clocks"):
        print(k, json.dumps(d.get(k))[:700])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference exit $?"; cut -c1-300 $O/bench_reference.json
for m in fp16x2 bf16x3; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_$m.csv \
    python bench.py --mode $m --steps 1 --warmup 0 --frames 16 --cpu_frames 0 --graph 0 --parity_frames 0 --extras 0 --ncu 0 > $O/ncu_list_$m.log 2>&1; echo "ncu list $m exit $?"
done
