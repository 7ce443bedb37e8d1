#!/bin/bash
# round-2 final validation: every -m gpu test (no -x: all failures at once)
set -x
mkdir -p gpurun_out/r2z
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 420 python -m pytest tests -q -m gpu --tb=line -p no:cacheprovider > gpurun_out/r2z/tests_gpu.log 2>&1
echo "rc_tests=$?" >> gpurun_out/r2z/tests_gpu.log
grep -E "passed|failed|FAILED|Error|rc_tests" gpurun_out/r2z/tests_gpu.log | tail -30
