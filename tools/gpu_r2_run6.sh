#!/usr/bin/env bash
# round 2, GPU visit 6: the whole GPU suite (new system golden, refinement mode 2, matcher tail fix)
set +e
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SKIP_BENCH=1 SKIP_NCU=1 SKIP_AB=1 bash tools/gpu_check.sh 2>&1 | grep -E "passed|failed|rc=|Error|error|worst" | head -40
timeout 200 python -m pytest tests/test_gpu_system.py -m gpu -x -q -s -p no:cacheprovider 2>&1 | grep -E "worst|passed|failed" | head
timeout 200 python tools/gpu_knn_mma_check.py 0 time 0 2>&1 | tail -4
