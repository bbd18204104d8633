#!/usr/bin/env bash
# round 2, GPU visit 15: checksums of the 8 streams (scene per pair of ranks), loop-closure detect time with 32 hypotheses, sanity
set +e
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python bench.py --print-checksums 2>&1 | tail -2
python tools/gpu_lc_bench.py 2>&1 | tail -3
SKIP_BENCH=1 SKIP_NCU=1 SKIP_AB=1 FILES="test_gpu_ba test_gpu_init test_gpu_loopclosure" bash tools/gpu_check.sh 2>&1 | grep -E "passed|failed|rc=|Error|error|assert" | head -30
