#!/bin/bash
# N-GPU bench of BASELINE configs[4] as written (8 fixed sequences, strong scaling) + the weak-scaling block + the bit-identity check.
# usage (under gpurun --gpus N): bash tools/gpu_multi.sh N
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out; O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo "build failed"; tail -5 $O/build.log; exit 1; }
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench N=$N exit $?"; tail -c 600 $O/bench_n$N.err
python - <<PY
import json
try:
    d = json.load(open("$O/bench_n$N.json"))
    for k in ("value", "ms_per_step", "n_gpus", "e2e", "weak", "collective_ms", "gather_bit_identical", "gather_check", "roofline_whole_step", "clocks"):
        print(k, json.dumps(d.get(k))[:400])
    print(d["config"]["workload"][:200])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/bench_reference_n$N.json 2> $O/bench_reference_n$N.err; echo "reference N=$N exit $?"; cut -c1-900 $O/bench_reference_n$N.json | grep -o '"value": [0-9.]*\|"cores": [0-9]*' | head -4
