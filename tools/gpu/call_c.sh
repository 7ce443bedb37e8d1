#!/bin/bash
set -x
mkdir -p gpurun_out/r2i
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu --tb=short > gpurun_out/r2i/tests_ops.log 2>&1
echo "rc_ops=$?" >> gpurun_out/r2i/tests_ops.log
tail -8 gpurun_out/r2i/tests_ops.log
timeout 120 python tools/bench_small.py > gpurun_out/r2i/bench_small.log 2>&1
cat gpurun_out/r2i/bench_small.log
timeout 200 python tools/step_ab.py "-" "-" > gpurun_out/r2i/step_ab.log 2>&1
tail -3 gpurun_out/r2i/step_ab.log
AB_KIND=sample timeout 100 python tools/step_ab.py "-" "-" > gpurun_out/r2i/sample_ab.log 2>&1
tail -3 gpurun_out/r2i/sample_ab.log
timeout 300 python -m pytest tests/test_gpu_model.py -q -m gpu -x -k "parity or reproducible" --tb=short > gpurun_out/r2i/tests_model.log 2>&1
echo "rc_model=$?" >> gpurun_out/r2i/tests_model.log
tail -4 gpurun_out/r2i/tests_model.log
