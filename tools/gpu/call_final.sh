#!/bin/bash
# end of round 2: the bench line (b200 arm) and the tests whose bounds changed last
set -x
mkdir -p gpurun_out/r2z
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
BENCH_WATCHDOG=200 timeout 240 python bench.py > gpurun_out/r2z/bench_final.json 2> gpurun_out/r2z/bench_final.err
echo "rc_bench=$?"
cat gpurun_out/r2z/bench_final.json | cut -c1-1500
tail -3 gpurun_out/r2z/bench_final.err
timeout 200 python -m pytest tests/test_gpu_model.py tests/test_gpu_conv_tc.py -q -m gpu -k "full_size and probunet or fused_statistics" --tb=line -p no:cacheprovider 2>&1 | tail -4
