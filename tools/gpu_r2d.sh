#!/bin/bash
# Round 2, call D: re-run of the GPU suite after the test rework (bf16 yardstick, training), drop-in eval breakdown.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo "build failed"; tail -5 $O/build.log; exit 1; }
timeout 1200 python -m pytest tests -q -m gpu -s > $O/pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; grep -E "passed|failed|error" $O/pytest_gpu.log | tail -n 5
grep -E "^bf16|bf16 |gradient errors|step [12]:|training losses|full-resolution training|FAILED|Error|^E  " $O/pytest_gpu.log | cut -c1-600 | head -n 70
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -n 5 $O/smoke.log | cut -c1-400
for c in 32 64 128; do
  timeout 300 python tools/dropin_eval_bench.py --frames 1024 --chunk $c > $O/dropin_c$c.json 2> $O/dropin_c$c.err; echo "dropin chunk $c exit $?"; cut -c1-900 $O/dropin_c$c.json; tail -n 2 $O/dropin_c$c.err
done
timeout 300 python tools/dropin_eval_bench.py --frames 256 --chunk 32 > $O/dropin_256.json 2> $O/dropin_256.err; cut -c1-900 $O/dropin_256.json
timeout 900 python bench.py --steps 5 --warmup 3 --dropin_frames 0 --parity_frames 0 --ncu 0 --extras 0 --cpu_frames 0 > $O/bench_quick.json 2> $O/bench_quick.err; echo "bench exit $?"; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_quick.json")); s = d["stages_ms"]
    print("quick: %.1f fps  e2e %.1f  sc %.3f tw %.3f of %.3f  clk %s" % (d["value"], d["e2e"]["value"], s["scoordnet_per_batch"], s["tower_per_batch"], s["oflownet_per_pair_batch"], d["clocks"].get("sm_mhz")))
except Exception as e:
    print("bench parse failed", e)
PY
