#!/bin/bash
# Round 2, call Q: CTA-pair plan of the limb modes (fp16x2 / bf16x3): parity suites + A/B (KFN_TC_PAIR_X3).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo "build failed"; exit 1; }
timeout 900 python -m pytest tests/test_gpu_networks.py tests/test_gpu_properties.py -q -m gpu -s > $O/pytest_q.log 2>&1; echo "pytest (networks, properties) exit $?"; grep -E "passed|failed|^E  |FAILED|^frame " $O/pytest_q.log | cut -c1-300 | tail -n 10
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -n 4 $O/smoke.log | cut -c1-200
q() { name=$1; shift; timeout 400 python bench.py --steps 3 --warmup 3 --cpu_frames 0 --parity_frames 0 --extras 0 --dropin_frames 0 --ncu 0 "$@" > $O/bench_$name.json 2> $O/bench_$name.err
      python - <<PY
import json
try:
    d = json.load(open("$O/bench_$name.json")); s = d["stages_ms"]
    print("$name: %.1f fps  e2e %.1f  sc %.3f tw %.3f of %.3f  clk %s" % (d["value"], d["e2e"]["value"], s["scoordnet_per_batch"], s["tower_per_batch"], s["oflownet_per_pair_batch"], d["clocks"].get("sm_mhz")))
except Exception as e:
    print("$name: bench failed", e); print(open("$O/bench_$name.err").read()[-500:])
PY
}
q fp16x2_pair --mode fp16x2
KFN_TC_PAIR_X3=0 q fp16x2_onecta --mode fp16x2
q bf16x3_pair --mode bf16x3
KFN_TC_PAIR_X3=0 q bf16x3_onecta --mode bf16x3
q fp16x2_pair2 --mode fp16x2
