#!/bin/bash
set -x
mkdir -p gpurun_out/r2n
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 300 python -m pytest tests/test_gpu_model.py tests/test_gpu_layers.py tests/test_input_pipeline.py tests/test_metrics.py -q -m gpu --tb=line -p no:cacheprovider > gpurun_out/r2n/tests_model.log 2>&1
echo "rc_model=$?" >> gpurun_out/r2n/tests_model.log
grep -E "passed|failed|FAILED|rc_model" gpurun_out/r2n/tests_model.log | tail -12
