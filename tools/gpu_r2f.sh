#!/bin/bash
# Round 2, call F: full GPU suite, smoke, the full default bench line (batch 32) and the reference arm -- the state the driver will see.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo "build failed"; tail -5 $O/build.log; exit 1; }
timeout 1200 python -m pytest tests -q -m gpu -s > $O/pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; grep -E "passed|failed|error" $O/pytest_gpu.log | tail -n 3
grep -E "^E  |FAILED|Adam:|gradient errors vs fp64 autograd \(rms-relative\): median|training losses|full-resolution" $O/pytest_gpu.log | cut -c1-400 | head -n 30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -n 4 $O/smoke.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; tail -c 300 $O/bench.err; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench.json"))
    for k in ("value", "ms_per_step", "e2e", "gpu_launches", "roofline", "roofline_whole_step", "stages_ms", "parity_mode", "configs4_on_one_gpu", "dropin_eval_from_files", "cpu_baseline", "clocks"):
        print(k, json.dumps(d.get(k))[:700])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference exit $?"; cut -c1-200 $O/bench_reference.json
timeout 300 python bench.py --steps 10 --warmup 3 --frames_per_batch 16 --pairs_per_batch 16 --cpu_frames 0 --parity_frames 0 --extras 0 --ncu 0 --dropin_frames 0 > $O/bench_b16.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_b16.json')); print('batch16: %.1f fps e2e %.1f clk %s' % (d['value'], d['e2e']['value'], d['clocks'].get('sm_mhz')))"
timeout 300 python bench.py --steps 10 --warmup 3 --cpu_frames 0 --parity_frames 0 --extras 0 --ncu 0 --dropin_frames 0 > $O/bench_b32.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_b32.json')); print('batch32: %.1f fps e2e %.1f clk %s' % (d['value'], d['e2e']['value'], d['clocks'].get('sm_mhz')))"
