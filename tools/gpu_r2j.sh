#!/bin/bash
# Round 2, call J: SCoordNet batch = 71 frames (integer number of waves on 148 SMs) with the tower in sub-batches; new property tests.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo "build failed"; tail -5 $O/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_properties.py tests/test_gpu_networks.py tests/test_gpu_eval_dropin.py -q -m gpu > $O/pytest_j.log 2>&1; echo "pytest (properties, networks, dropin) exit $?"; grep -E "passed|failed|^E  |FAILED" $O/pytest_j.log | cut -c1-300 | tail -n 12
q() { name=$1; shift; timeout 300 python bench.py --steps 10 --warmup 3 --cpu_frames 0 --parity_frames 0 --extras 0 --ncu 0 --dropin_frames 0 "$@" > $O/bench_$name.json 2> $O/bench_$name.err
      python - <<PY
import json
try:
    d = json.load(open("$O/bench_$name.json")); s = d["stages_ms"]
    print("$name: %.1f fps  e2e %.1f  sc %.3f tw %.3f of %.3f  clk %s" % (d["value"], d["e2e"]["value"], s["scoordnet_per_batch"], s["tower_per_batch"], s["oflownet_per_pair_batch"], d["clocks"].get("sm_mhz")))
except Exception as e:
    print("$name: bench failed", e)
PY
}
q f71p32
q f32p32 --frames_per_batch 32
q f71p32_again
q f63p32 --frames_per_batch 63
q f71p16 --pairs_per_batch 16
q f128p32 --frames_per_batch 128
q f16p16 --frames_per_batch 16 --pairs_per_batch 16
