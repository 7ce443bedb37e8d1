#!/bin/bash
# round-2 late validation: new kernels (two-launch batch-norm backward, sample moments, input pipeline) + whole-step A/B
set -x
mkdir -p gpurun_out/r2b
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 420 python -m pytest tests/test_input_pipeline.py tests/test_gpu_conv_tc.py -x -q -m gpu > gpurun_out/r2b/tests_new.log 2>&1
echo "rc_new=$?" >> gpurun_out/r2b/tests_new.log
tail -5 gpurun_out/r2b/tests_new.log
timeout 260 python tools/step_ab.py "-" "PHS_STATS_MIN_HW=512" "PHS_STATS_MIN_HW=2048" "PHS_STATS_MIN_HW=8192" "PHS_STATS_MIN_HW=2048 PHS_HALO_PAIR=1" "-" > gpurun_out/r2b/step_ab.log 2>&1
cat gpurun_out/r2b/step_ab.log | tail -6
timeout 300 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "reproducible" > gpurun_out/r2b/tests_model.log 2>&1
echo "rc_model=$?" >> gpurun_out/r2b/tests_model.log
tail -5 gpurun_out/r2b/tests_model.log
PHS_STATS_MIN_HW=2048 timeout 300 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "parity or reproducible or fast_mode" > gpurun_out/r2b/tests_split.log 2>&1; echo "rc_split=$?" >> gpurun_out/r2b/tests_split.log; tail -4 gpurun_out/r2b/tests_split.log
