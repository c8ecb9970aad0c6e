#!/bin/bash
# First GPU call of the next round: re-validate HEAD, then try the opt-in fused conv1a+conv1b kernel
# (kfnet_b200/csrc/conv1ab_fused.cu, KFN_TC_FUSE_CONV1=1; written at the end of round 1 without GPU time left).
# The kernel's mbarrier waits trap after ~2 s instead of hanging; every step also runs under `timeout`.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo "build failed"; tail -5 $O/build.log; exit 1; }
timeout 300 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest gpu (defaults) exit $?"; tail -n 4 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 $O/smoke.log
KFN_TC_FUSE_CONV1=1 timeout 120 python -m pytest tests/test_gpu_conv_layers.py -q -m gpu -k fused_conv1ab > $O/t_fused.log 2>&1
frc=$?; echo "fused conv1ab layer test exit $frc"; tail -n 12 $O/t_fused.log
b() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 5 --warmup 3 --cpu_frames 0 --parity_frames 0 --extras 0 --ncu 0 > $O/bench_$name.json 2> $O/bench_$name.err
      python - <<PY
import json
try:
    d = json.load(open("$O/bench_$name.json")); s = d["stages_ms"]
    print("$name: %.1f fps  e2e %.1f  sc %.3f tw %.3f of %.3f  clk %s" % (d["value"], d["e2e"]["value"], s["scoordnet_per_batch"], s["tower_per_batch"], s["oflownet_per_pair_batch"], d["clocks"].get("sm_mhz")))
except Exception as e:
    print("$name: bench failed", e)
PY
}
b default X=1
if [ $frc -eq 0 ]; then
  KFN_TC_FUSE_CONV1=1 timeout 300 python -m pytest tests/test_gpu_networks.py tests/test_gpu_properties.py tests/test_gpu_eval_dropin.py -q -m gpu > $O/pytest_fused_paths.log 2>&1
  echo "network/property/drop-in tests with the fused kernel exit $?"; tail -n 3 $O/pytest_fused_paths.log
  b fused KFN_TC_FUSE_CONV1=1
  b default2 X=1
  KFN_TC_FUSE_CONV1=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_fused.csv \
      python bench.py --steps 1 --warmup 0 --frames 16 --cpu_frames 0 --graph 0 --parity_frames 0 --extras 0 --ncu 0 > $O/ncu_list_fused.log 2>&1; echo "ncu list (fused) exit $?"
  KFN_TC_FUSE_CONV1=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv1ab -c 1 -f -o $O/prof_conv1ab \
      python bench.py --steps 1 --warmup 0 --frames 16 --cpu_frames 0 --graph 0 --parity_frames 0 --extras 0 --ncu 0 > $O/ncu_full_fused.log 2>&1; echo "ncu full (fused) exit $?"
fi
# second opt-in variant: U-Net conv0 forming the difference volume in its producer warps (KFN_TC_DIFF_FUSED=1)
timeout 100 python tools/ab_compare.py $O/ab_default.npz > $O/ab_default.log 2>&1
KFN_TC_DIFF_FUSED=1 timeout 100 python tools/ab_compare.py $O/ab_diff.npz $O/ab_default.npz > $O/ab_diff.log 2>&1
drc=$?; echo "fused difference volume: bit-compare vs default exit $drc"; tail -n 8 $O/ab_diff.log
if [ $drc -eq 0 ]; then
  KFN_TC_DIFF_FUSED=1 timeout 300 python -m pytest tests/test_gpu_networks.py tests/test_gpu_properties.py tests/test_gpu_eval_dropin.py -q -m gpu > $O/pytest_diff_paths.log 2>&1
  echo "network/property/drop-in tests with the fused difference volume exit $?"; tail -n 3 $O/pytest_diff_paths.log
  b diff_fused KFN_TC_DIFF_FUSED=1
  b default3 X=1
fi
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 1 --warmup 0 --frames 16 --cpu_frames 0 --graph 0 --parity_frames 0 --extras 0 --ncu 0 > $O/ncu_list.log 2>&1; echo "ncu list exit $?"
