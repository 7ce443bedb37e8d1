#!/bin/bash
set -x
mkdir -p gpurun_out/r2m
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 200 python -m pytest tests/test_gpu_ops.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r2m/tests_ops.log 2>&1
echo "rc_ops=$?" >> gpurun_out/r2m/tests_ops.log
tail -12 gpurun_out/r2m/tests_ops.log | cut -c1-250
AB_WATCHDOG=100 timeout 150 python tools/step_ab.py "PHS_NO_WGRAD_SMALL3=1" "-" "PHS_NO_WGRAD_SMALL3=1" "-" > gpurun_out/r2m/step_ab.log 2>&1
tail -6 gpurun_out/r2m/step_ab.log
