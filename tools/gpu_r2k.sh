#!/bin/bash
# Round 2, call K: interleaved tile assignment (KFN_TC_STRIDED) -- parity at full size, A/B, DRAM bytes per SCoordNet launch both ways.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo "build failed"; tail -5 $O/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_networks.py tests/test_gpu_properties.py tests/test_gpu_conv_layers.py -q -m gpu > $O/pytest_k.log 2>&1; echo "pytest (networks, properties, layers) exit $?"; grep -E "passed|failed|^E  |FAILED" $O/pytest_k.log | cut -c1-300 | tail -n 12
q() { name=$1; shift; timeout 400 python bench.py --steps 10 --warmup 3 --cpu_frames 0 --parity_frames 0 --extras 0 --dropin_frames 0 "$@" > $O/bench_$name.json 2> $O/bench_$name.err
      python - <<PY
import json
try:
    d = json.load(open("$O/bench_$name.json")); s = d["stages_ms"]; t = d["roofline"].get("traffic") or {}
    print("$name: %.1f fps  e2e %.1f  sc %.3f tw %.3f of %.3f  clk %s  dram/batch %.2f GB %s" % (d["value"], d["e2e"]["value"], s["scoordnet_per_batch"], s["tower_per_batch"], s["oflownet_per_pair_batch"], d["clocks"].get("sm_mhz"), (t.get("bytes") or 0) / 1e9, [round(x / 1e6) for x in (t.get("per_launch_bytes") or [])]))
except Exception as e:
    print("$name: bench failed", e)
PY
}
q strided
KFN_TC_STRIDED=0 q contiguous
q strided2 --ncu 0
KFN_TC_STRIDED=0 q contiguous2 --ncu 0
q strided_b16 --ncu 0 --frames_per_batch 16 --pairs_per_batch 16
q strided_b64 --ncu 0 --frames_per_batch 64
