#!/bin/bash
# Round 2, call O: pairs-per-batch A/B with the CTA-pair build (the OFlowNet share of the step grew to 23 %).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo "build failed"; exit 1; }
q() { name=$1; shift; timeout 400 python bench.py --steps 10 --warmup 3 --cpu_frames 0 --parity_frames 0 --extras 0 --dropin_frames 0 --ncu 0 "$@" > $O/bench_$name.json 2> $O/bench_$name.err
      python - <<PY
import json
try:
    d = json.load(open("$O/bench_$name.json")); s = d["stages_ms"]
    print("$name: %.1f fps  e2e %.1f  sc %.3f tw %.3f of %.3f  clk %s" % (d["value"], d["e2e"]["value"], s["scoordnet_per_batch"], s["tower_per_batch"], s["oflownet_per_pair_batch"], d["clocks"].get("sm_mhz")))
except Exception as e:
    print("$name: bench failed", e)
PY
}
q p32
q p16 --pairs_per_batch 16
q p8 --pairs_per_batch 8
q p32b
q p16b --pairs_per_batch 16
q p24 --pairs_per_batch 24
q f16p16 --frames_per_batch 16 --pairs_per_batch 16
q f48p16 --frames_per_batch 48 --pairs_per_batch 16
