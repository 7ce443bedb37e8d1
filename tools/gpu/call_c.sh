#!/bin/bash
set -x
mkdir -p gpurun_out/r2k
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
AB_WATCHDOG=45 timeout 60 python tools/step_ab.py "-" > gpurun_out/r2k/step_ab_new.log 2>&1
tail -25 gpurun_out/r2k/step_ab_new.log
AB_WATCHDOG=45 timeout 60 python tools/step_ab.py "PHS_NORM_RAW=0" > gpurun_out/r2k/step_ab_old.log 2>&1
tail -25 gpurun_out/r2k/step_ab_old.log
