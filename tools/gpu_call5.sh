set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_ops.py tests/test_gpu_layers.py -m gpu -q -x 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -4
for pdl in 1 0; do
PHS_PDL=$pdl timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2/bench3_pdl$pdl.json 2> gpurun_out/r2/bench3_pdl$pdl.err
python -c "
import json
d=json.load(open('gpurun_out/r2/bench3_pdl$pdl.json'))
print('PDL=$pdl ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'sampling', d['sampling']['value'], 'loss', d['loss'])
"
tail -2 gpurun_out/r2/bench3_pdl$pdl.err
done
SWEEP_CFGS="2/-/-/-" timeout 300 python tools/sweep_halo.py trace > gpurun_out/r2/trace1.txt 2>&1
grep -A7 "128x128  32->32  cfg" gpurun_out/r2/trace1.txt | cut -c1-400
