set -x
mkdir -p gpurun_out/r2
SWEEP_CFGS="2/-/-/-,2/4/-/-,2/2/-/-,2/1/-/-,1/-/-/-,1/2/-/-,1/1/-/-" timeout 600 python tools/sweep_halo.py > gpurun_out/r2/sweep_S.txt 2>&1
cat gpurun_out/r2/sweep_S.txt
