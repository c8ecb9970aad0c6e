#!/bin/bash
# One-shot validation of the epilogue/first-layer kernel variants on one B200 (profiles/r01h_*).
# Each variant has an env switch (KFN_TC_WIDE, KFN_TC_IM2COL_FUSED, KFN_TC_GROUP2); the run sets them explicitly.  A variant whose own
# tests fail is switched OFF for the rest of the run, so the "final" configuration that gets the full test suite, the
# bench and the ncu captures is always one that passed; the all-legacy configuration is run beside it for the A/B.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo "build failed"; tail -5 $O/build.log; exit 1; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1

CFG="KFN_TC_WIDE=1 KFN_TC_IM2COL_FUSED=1 KFN_TC_GROUP2=1"
off() { CFG="${CFG/$1=1/$1=0}"; }
pt() { env $CFG timeout 300 python -m pytest -q -m gpu "$@"; }

pt tests/test_gpu_conv_layers.py -k "wide_layout" > $O/t1_wide.log 2>&1
if [ $? -ne 0 ]; then off KFN_TC_WIDE; echo "WIDE failed -> off"; tail -n 15 $O/t1_wide.log; else echo "WIDE ok"; fi
pt tests/test_gpu_conv_layers.py -k "first_layer" > $O/t2_first.log 2>&1
if [ $? -ne 0 ]; then off KFN_TC_IM2COL_FUSED; echo "IM2COL_FUSED failed -> off"; tail -n 15 $O/t2_first.log; else echo "IM2COL_FUSED ok"; fi
pt tests/test_gpu_conv_layers.py --deselect tests/test_gpu_conv_layers.py::test_first_layer_fused_from_u8 \
   --deselect tests/test_gpu_conv_layers.py::test_wide_layout_conv1b_full_frame > $O/t3_layers.log 2>&1
if [ $? -ne 0 ]; then off KFN_TC_GROUP2; echo "layer tests failed with GROUP2 -> off"; tail -n 15 $O/t3_layers.log
  pt tests/test_gpu_conv_layers.py --deselect tests/test_gpu_conv_layers.py::test_first_layer_fused_from_u8 \
     --deselect tests/test_gpu_conv_layers.py::test_wide_layout_conv1b_full_frame > $O/t3b_layers.log 2>&1; echo "layer tests without GROUP2: exit $?"
else echo "GROUP2 ok"; fi
echo "FINAL_CFG=[$CFG]" | tee $O/final_cfg.txt

LEG="KFN_TC_WIDE=0 KFN_TC_IM2COL_FUSED=0 KFN_TC_GROUP2=0"
env $CFG timeout 400 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest gpu (final cfg) exit $?"; tail -n 4 $O/pytest_gpu.log
env $LEG timeout 400 python -m pytest tests -q -m gpu > $O/pytest_gpu_legacy.log 2>&1; echo "pytest gpu (legacy cfg) exit $?"; tail -n 4 $O/pytest_gpu_legacy.log
env $CFG python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 $O/smoke.log

b() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 5 --warmup 3 --cpu_frames 0 > $O/bench_$name.json 2> $O/bench_$name.err
      python - <<EOF
import json
try:
    d = json.load(open("$O/bench_$name.json")); s = d["stages_ms"]
    print("$name: %.1f fps  e2e %.1f  sc %.3f tw %.3f of %.3f  clk %s" % (d["value"], d["e2e"]["value"], s["scoordnet_per_batch"], s["tower_per_batch"], s["oflownet_per_pair_batch"], d["clocks"].get("sm_mhz")))
except Exception as e:
    print("$name: bench failed", e)
EOF
}
b final $CFG
b legacy $LEG
b no_wide $CFG KFN_TC_WIDE=0
b no_im2col $CFG KFN_TC_IM2COL_FUSED=0
b no_group2 $CFG KFN_TC_GROUP2=0
env $CFG timeout 300 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench (final cfg, with cpu baseline) exit $?"; cut -c1-200 $O/bench.json

env $CFG timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 1 --warmup 0 --frames 16 --cpu_frames 0 --graph 0 > $O/ncu_list.log 2>&1; echo "ncu list exit $?"
env $CFG timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 0 -c 2 -f -o $O/prof_first \
    python bench.py --steps 1 --warmup 0 --frames 16 --cpu_frames 0 --graph 0 > $O/ncu_full0.log 2>&1; echo "ncu first+conv1b exit $?"
env $CFG timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 7 -c 1 -f -o $O/prof_conv_tc \
    python bench.py --steps 1 --warmup 0 --frames 16 --cpu_frames 0 --graph 0 > $O/ncu_full1.log 2>&1; echo "ncu conv4b exit $?"
