#!/bin/bash
# batch-size sweep of the bench (one gpurun call)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for cfg in "16 16" "8 8" "32 16" "16 32"; do
  set -- $cfg
  python bench.py --steps 2 --warmup 2 --cpu_frames 0 --frames_per_batch $1 --pairs_per_batch $2 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read()); print('fpb/ppb', b['config']['frames_per_batch'], b['config']['pairs_per_batch'], 'fps', round(b['value'],1), 'e2e', round(b['e2e']['value'],1), {k:round(v,3) for k,v in b['stages_ms'].items() if 'per' in k and 'GB' not in k}, 'roofline', round(b['roofline']['frac'],3))"
done
