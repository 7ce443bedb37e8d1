#!/bin/bash
# conv -> norm -> ReLU -> conv fusion: whole-step equivalence (all cases, with the per-tensor diagnostic), in-step A/B by layer class
set -x
mkdir -p gpurun_out/r2d
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 100 python -m pytest tests/test_gpu_fused_norm.py -x -q -m gpu -k "remat" > gpurun_out/r2d/tests_kernel2.log 2>&1
echo "rc_kernel=$?" >> gpurun_out/r2d/tests_kernel2.log
tail -3 gpurun_out/r2d/tests_kernel2.log
timeout 400 python -m pytest tests/test_gpu_fused_norm.py -q -m gpu -s -k "same_with" --tb=line > gpurun_out/r2d/tests_step.log 2>&1
echo "rc_step=$?" >> gpurun_out/r2d/tests_step.log
grep -v "^$" gpurun_out/r2d/tests_step.log | tail -70
timeout 300 python tools/step_ab.py "-" "PHS_FUSE_NORM=1" "PHS_FUSE_NORM=1 PHS_FUSE_MINCIN=128" "PHS_FUSE_NORM=1 PHS_FUSE_MINCIN=64" "PHS_FUSE_NORM=1 PHS_FUSE_MAXHW=4096" "PHS_FUSE_NORM=1 PHS_FUSE_MAXHW=1024" "-" > gpurun_out/r2d/step_ab2.log 2>&1
cat gpurun_out/r2d/step_ab2.log | tail -8
