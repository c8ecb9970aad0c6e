#!/bin/bash
# Round 2, call C: bf16 tests against the storage model, streaming drop-in eval, full bench, batch-size A/B.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo "build failed"; tail -5 $O/build.log; exit 1; }
timeout 900 python -m pytest tests -q -m gpu -s > $O/pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; grep -E "passed|failed|error" $O/pytest_gpu.log | tail -n 5
grep -E "bf16|storage|FAILED|Error|assert" $O/pytest_gpu.log | cut -c1-400 | head -n 60
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -n 5 $O/smoke.log | cut -c1-400
timeout 300 python tools/dropin_eval_bench.py --frames 256 > $O/dropin.json 2> $O/dropin.err; echo "dropin exit $?"; cat $O/dropin.json | cut -c1-500; tail -n 3 $O/dropin.err
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; tail -c 400 $O/bench.err; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench.json"))
    for k in ("value", "ms_per_step", "e2e", "parity_mode", "dropin_eval_from_files", "cpu_baseline"):
        print(k, json.dumps(d.get(k))[:900])
except Exception as e:
    print("bench parse failed", e)
PY
q() { name=$1; shift; timeout 300 python bench.py --steps 5 --warmup 3 --cpu_frames 0 --parity_frames 0 --extras 0 --ncu 0 --dropin_frames 0 "$@" > $O/bench_$name.json 2> $O/bench_$name.err
      python - <<PY
import json
try:
    d = json.load(open("$O/bench_$name.json")); s = d["stages_ms"]
    print("$name: %.1f fps  e2e %.1f  sc %.3f tw %.3f of %.3f  clk %s" % (d["value"], d["e2e"]["value"], s["scoordnet_per_batch"], s["tower_per_batch"], s["oflownet_per_pair_batch"], d["clocks"].get("sm_mhz")))
except Exception as e:
    print("$name: bench failed", e)
PY
}
q b16
q b32 --frames_per_batch 32 --pairs_per_batch 32
q b16p32 --pairs_per_batch 32
q b16b
# error-compensated modes: halo tiles on/off
q fp16x2_halo --mode fp16x2 --steps 2
KFN_TC_HALO_X3=0 q fp16x2_nohalo --mode fp16x2 --steps 2
q bf16x3_halo --mode bf16x3 --steps 2
