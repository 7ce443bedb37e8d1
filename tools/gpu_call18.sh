set -x
timeout 200 python tools/step_ab.py "PHS_HALO_PAIR=1" "PHS_HALO_PAIR=2" "PHS_HALO_PAIR=2 PHS_HALO_G=64" "PHS_HALO_PAIR=0" "PHS_HALO_PAIR=2 PHS_HALO_PAIR_MINCIN=64" 2>&1 | grep "ms/step"
