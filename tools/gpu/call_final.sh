#!/bin/bash
# end of round 2: the bench line (b200 arm)
set -x
mkdir -p gpurun_out/r2z
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
BENCH_WATCHDOG=150 timeout 180 python bench.py > gpurun_out/r2z/bench_final.json 2> gpurun_out/r2z/bench_final.err
echo "rc_bench=$?"
cut -c1-1200 gpurun_out/r2z/bench_final.json
tail -2 gpurun_out/r2z/bench_final.err
