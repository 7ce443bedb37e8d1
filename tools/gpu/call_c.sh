#!/bin/bash
set -x
mkdir -p gpurun_out/r2f
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
DIAG_ENV_A="PHS_FUSE_NORM=0" DIAG_ENV_B="PHS_FUSE_NORM=1" timeout 200 python tools/diag_determinism.py phiseg_7_5 128 4 > gpurun_out/r2f/diag_fusion.log 2>&1
cut -c1-400 gpurun_out/r2f/diag_fusion.log | tail -45
timeout 300 python -m pytest tests/test_gpu_fused_norm.py tests/test_gpu_model.py -q -m gpu -s -k "folded or (probunet and (sampling or batched))" --tb=short > gpurun_out/r2f/tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2f/tests.log
grep -E "passed|failed|FAILED|fast sampling|folded|Error|assert " gpurun_out/r2f/tests.log | tail -12
