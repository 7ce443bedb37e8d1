set -x
mkdir -p gpurun_out/r2
python tools/diag_determinism.py phiseg_7_5 128 8 > gpurun_out/r2/diag_lanes.txt 2>&1
PHS_NO_LANES=1 python tools/diag_determinism.py phiseg_7_5 128 8 > gpurun_out/r2/diag_nolanes.txt 2>&1
PHS_NO_WLANE=1 python tools/diag_determinism.py phiseg_7_5 128 8 > gpurun_out/r2/diag_nowlane.txt 2>&1
head -50 gpurun_out/r2/diag_lanes.txt; head -30 gpurun_out/r2/diag_nolanes.txt; head -30 gpurun_out/r2/diag_nowlane.txt
timeout 300 compute-sanitizer --tool initcheck --print-limit 8 python tools/sanitize_step.py phiseg_7_5 64 2 2>&1 | head -150 > gpurun_out/r2/san_initcheck.log
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | tail -150 > gpurun_out/r2/pytest2.log
grep -E "passed|failed|FAILED|reproducib|fast |sampling|dp-equiv" gpurun_out/r2/pytest2.log | head -60
du -sh gpurun_out
