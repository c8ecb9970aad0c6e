#!/bin/bash
# Round 2, call Y: resident weights for the small non-halo layers (KFN_TC_WRES): bit-compare, parity suites, A/B, launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo "build failed"; exit 1; }
KFN_TC_WRES=0 timeout 120 python tools/ab_compare.py $O/ab_old.npz > $O/ab_old.log 2>&1; echo "ab old exit $?"
timeout 120 python tools/ab_compare.py $O/ab_new.npz $O/ab_old.npz > $O/ab_new.log 2>&1
echo "resident weights: bit-compare vs staged exit $?"; tail -n 7 $O/ab_new.log | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_conv_layers.py tests/test_gpu_networks.py -q -m gpu -x > $O/pytest_y.log 2>&1; echo "pytest (layers, networks) exit $?"; grep -E "passed|failed|^E  |FAILED" $O/pytest_y.log | cut -c1-300 | tail -n 8
q() { name=$1; shift; timeout 300 python bench.py --steps 5 --warmup 3 --cpu_frames 0 --parity_frames 0 --extras 0 --dropin_frames 0 --ncu 0 "$@" > $O/bench_$name.json 2> $O/bench_$name.err
      python - <<PY
import json
try:
    d = json.load(open("$O/bench_$name.json")); s = d["stages_ms"]
    print("$name: %.1f fps  e2e %.1f  sc %.3f tw %.3f of %.3f  clk %s" % (d["value"], d["e2e"]["value"], s["scoordnet_per_batch"], s["tower_per_batch"], s["oflownet_per_pair_batch"], d["clocks"].get("sm_mhz")))
except Exception as e:
    print("$name: bench failed", e); print(open("$O/bench_$name.err").read()[-600:])
PY
}
q wres
KFN_TC_WRES=0 q staged
q wres2
KFN_TC_WRES=0 q staged2
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/launches.csv \
    python bench.py --steps 1 --warmup 0 --frames 16 --frames_per_batch 16 --pairs_per_batch 16 --cpu_frames 0 --graph 0 --parity_frames 0 --extras 0 --ncu 0 --dropin_frames 0 > $O/ncu_list.log 2>&1; echo "ncu list exit $?"
