#!/bin/bash
set -x
mkdir -p gpurun_out/r2g
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 500 python -m pytest tests/test_cli.py tests/test_gpu_model.py tests/test_gpu_fused_norm.py -q -m gpu -s -k "cli or detunet or session or same_with" --tb=short > gpurun_out/r2g/tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2g/tests.log
grep -E "passed|failed|FAILED|fast sampling|fusion|cosine|Error|assert |^E  " gpurun_out/r2g/tests.log | tail -40
