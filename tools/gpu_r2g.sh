#!/bin/bash
# Round 2, call G: fused-limb stages of the error-compensated modes (tests + A/B), batch-size A/B on one box, training loop on files.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo "build failed"; tail -5 $O/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_networks.py tests/test_gpu_training.py tests/test_gpu_eval_dropin.py tests/test_gpu_conv_layers.py -q -m gpu > $O/pytest_g.log 2>&1; echo "pytest (networks, training, dropin, layers) exit $?"; grep -E "passed|failed|^E  |FAILED" $O/pytest_g.log | cut -c1-300 | tail -n 12
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -n 4 $O/smoke.log | cut -c1-200
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
q f32p32
q f32p16 --pairs_per_batch 16
q f16p16 --frames_per_batch 16 --pairs_per_batch 16
q f32p32_again
q fp16x2_fused --mode fp16x2 --steps 3
KFN_TC_X3_FUSED=0 q fp16x2_unfused --mode fp16x2 --steps 3
q bf16x3_fused --mode bf16x3 --steps 3
KFN_TC_X3_FUSED=0 q bf16x3_unfused --mode bf16x3 --steps 3
