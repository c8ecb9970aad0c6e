#!/bin/bash
# Round 2, call V: scan cluster size sweep (launch-attribute cluster, next-frame input prefetch) + the scan / property tests.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo "build failed"; exit 1; }
for c in 4 5 6 8 3; do KFN_SCAN_CLUSTER_SIZE=$c timeout 120 python tools/exp_scan.py 2>&1 | tail -2; done
KFN_SCAN_CLUSTER=0 timeout 120 python tools/exp_scan.py 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_properties.py tests/test_gpu_networks.py -q -m gpu > $O/pytest_v.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|^E  |FAILED" $O/pytest_v.log | cut -c1-300 | tail -n 6
KFN_SCAN_CLUSTER_SIZE=5 timeout 600 python -m pytest tests/test_gpu_properties.py tests/test_gpu_networks.py -q -m gpu > $O/pytest_v5.log 2>&1; echo "pytest (cluster 5) exit $?"; grep -E "passed|failed|^E  |FAILED" $O/pytest_v5.log | cut -c1-300 | tail -n 6
