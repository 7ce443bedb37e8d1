set -x
timeout 150 python tools/step_ab.py "-" "PHS_NORM_BPS=2" "PHS_NORM_BPS=4" "PHS_NORM_BPS=6" "PHS_NORM_BPS=8" "PHS_FUSED_BN_BWD=1" "PHS_WGRAD_CTAS=1" "PHS_WGRAD_CTAS=2" 2>&1 | grep "ms/step"
