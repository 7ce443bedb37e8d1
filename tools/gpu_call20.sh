set -x
timeout 150 python tools/step_ab.py "-" "PHS_NORM_BPS=2" "PHS_NORM_BPS=4" "PHS_NORM_BPS=6" "PHS_NO_PAD=1" "PHS_NO_HALO=1" 2>&1 | grep "ms/step"
