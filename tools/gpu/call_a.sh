#!/bin/bash
# round-2 late validation: new kernels (two-launch batch-norm backward, sample moments, input pipeline) + whole-step A/B
set -x
mkdir -p gpurun_out/r2b
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 420 python -m pytest tests/test_gpu_ops.py tests/test_metrics.py tests/test_input_pipeline.py -x -q -m gpu > gpurun_out/r2b/tests_new.log 2>&1
echo "rc_new=$?" >> gpurun_out/r2b/tests_new.log
tail -5 gpurun_out/r2b/tests_new.log
timeout 200 python tools/step_ab.py "-" "PHS_BN_BWD3=1" "PHS_HALO_PAIR=1" "-" > gpurun_out/r2b/step_ab.log 2>&1
cat gpurun_out/r2b/step_ab.log | tail -6
timeout 300 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "parity or reproducible or predict_api" > gpurun_out/r2b/tests_model.log 2>&1
echo "rc_model=$?" >> gpurun_out/r2b/tests_model.log
tail -5 gpurun_out/r2b/tests_model.log
timeout 240 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b/bench.json 2> gpurun_out/r2b/bench.err
cat gpurun_out/r2b/bench.json; tail -3 gpurun_out/r2b/bench.err
