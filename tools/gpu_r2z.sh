#!/bin/bash
# Round 2, call Z (last): full GPU suite + smoke + the default bench line on the final build.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo "build failed"; exit 1; }
timeout 240 python -m pytest tests -q -m gpu -x -s > $O/pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; grep -E "passed|failed|^E  |FAILED" $O/pytest_gpu.log | cut -c1-300 | tail -n 6
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -n 4 $O/smoke.log | cut -c1-200
timeout 200 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench exit $?"
python - <<PY
import json
d = json.load(open("$O/bench.json"))
for k in ("value", "ms_per_step", "e2e", "gpu_launches", "roofline_whole_step", "stages_ms", "dropin_eval_from_files", "configs4_on_one_gpu", "clocks"):
    print(k, json.dumps(d.get(k))[:300])
print("roofline frac", d["roofline"]["frac"], "traffic ratio", (d["roofline"].get("traffic") or {}).get("ratio"))
for m, v in d.get("parity_mode", {}).items():
    if isinstance(v, dict) and "frames_per_s" in v: print("parity", m, v["frames_per_s"])
PY
