#!/bin/bash
# One gpurun call: build, then each GPU test file in its own process under a timeout (a trapped kernel kills
# only that process), logs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for f in "$@"; do
  name=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -q -s -m gpu --tb=short > "gpurun_out/$name.log" 2>&1
  r=$?
  echo "== $f -> exit $r"; tail -n 30 "gpurun_out/$name.log"
  [ $r -ne 0 ] && rc=$r
done
exit $rc
