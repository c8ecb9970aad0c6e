#!/bin/bash
# Round 2, call M: CTA-pair kernel (cta_group::2, KFN_TC_PAIR=1): bit-compare against the one-CTA path, parity suites, A/B.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo "build failed"; exit 1; }
timeout 120 python tools/ab_compare.py $O/ab_default.npz > $O/ab_default.log 2>&1; echo "ab default exit $?"
KFN_TC_PAIR=1 timeout 120 python tools/ab_compare.py $O/ab_pair.npz $O/ab_default.npz > $O/ab_pair.log 2>&1
prc=$?; echo "CTA-pair kernel: bit-compare vs default exit $prc"; tail -n 9 $O/ab_pair.log | cut -c1-200
if true; then
  KFN_TC_PAIR=1 timeout 600 python -m pytest tests/test_gpu_networks.py tests/test_gpu_properties.py tests/test_gpu_eval_dropin.py -q -m gpu > $O/pytest_pair.log 2>&1; echo "network / property / drop-in tests with the pair kernel exit $?"; grep -E "passed|failed|^E  " $O/pytest_pair.log | cut -c1-300 | tail -n 6
fi
q() { name=$1; shift; timeout 400 python bench.py --steps 10 --warmup 3 --cpu_frames 0 --parity_frames 0 --extras 0 --dropin_frames 0 --ncu 0 "$@" > $O/bench_$name.json 2> $O/bench_$name.err
      python - <<PY
import json
try:
    d = json.load(open("$O/bench_$name.json")); s = d["stages_ms"]
    print("$name: %.1f fps  e2e %.1f  sc %.3f tw %.3f of %.3f  clk %s" % (d["value"], d["e2e"]["value"], s["scoordnet_per_batch"], s["tower_per_batch"], s["oflownet_per_pair_batch"], d["clocks"].get("sm_mhz")))
except Exception as e:
    print("$name: bench failed", e); print(open("$O/bench_$name.err").read()[-600:])
PY
}
q one_cta
KFN_TC_PAIR=1 q pair
q one_cta2
KFN_TC_PAIR=1 q pair2
KFN_TC_PAIR=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/launches_pair.csv \
    python bench.py --steps 1 --warmup 0 --frames 16 --frames_per_batch 16 --pairs_per_batch 16 --cpu_frames 0 --graph 0 --parity_frames 0 --extras 0 --ncu 0 --dropin_frames 0 > $O/ncu_list_pair.log 2>&1; echo "ncu list (pair) exit $?"
