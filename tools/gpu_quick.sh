#!/bin/bash
# last short validation of round 1: lean co-resident kernel variant (KFN_TC_LEAN_X2) -- layer + network parity, A/B bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 100 python -m pytest tests/test_gpu_conv_layers.py tests/test_gpu_networks.py tests/test_gpu_eval_dropin.py -q -m gpu > $O/pytest_quick.log 2>&1; echo "pytest exit $?"; tail -n 3 $O/pytest_quick.log
for v in 1 0; do
  KFN_TC_LEAN_X2=$v timeout 60 python bench.py --steps 5 --warmup 3 --cpu_frames 0 > $O/bench_lean$v.json 2> $O/bench_lean$v.err
  python -c "
import json; d=json.load(open('$O/bench_lean$v.json')); s=d['stages_ms']
print('lean=$v: %.1f fps e2e %.1f sc %.3f tw %.3f of %.3f clk %s' % (d['value'], d['e2e']['value'], s['scoordnet_per_batch'], s['tower_per_batch'], s['oflownet_per_pair_batch'], d['clocks'].get('sm_mhz')))"
done
