set -x
timeout 600 python tools/step_ab.py "-" "PHS_HALO_CTAS=2" "PHS_HALO_NA=3" "PHS_HALO_G=32" "PHS_WLANES=1" "PHS_HALO_PAIR=1" "-" 2>&1 | grep "ms/step"
