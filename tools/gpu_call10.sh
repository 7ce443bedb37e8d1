set -x
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -s -k "parity_tc or reproducible or fast_mode" 2>&1 | grep -E "passed|failed|FAILED|worst relative|logit diff|Error|assert |reproducibility" | cut -c1-200 | head -30
SH="64,128,128,128,128;64,64,64,192,192;64,128,128,64,128;64,32,32,128,128;64,16,16,192,192;64,128,128,32,192;64,128,128,32,32;64,64,64,64,64"
PHS_HALO_SW=0 SWEEP_SHAPES="$SH" SWEEP_CFGS="2/-/-/-" timeout 300 python tools/sweep_halo.py > gpurun_out/r2/sweep_sw0.txt 2>&1
SWEEP_SHAPES="$SH" SWEEP_CFGS="2/-/-/-" timeout 300 python tools/sweep_halo.py > gpurun_out/r2/sweep_sw1.txt 2>&1
PHS_HALO_PAIR=1 SWEEP_SHAPES="$SH" SWEEP_CFGS="2/-/-/-" timeout 300 python tools/sweep_halo.py > gpurun_out/r2/sweep_sw1_pair.txt 2>&1
paste gpurun_out/r2/sweep_sw0.txt gpurun_out/r2/sweep_sw1.txt gpurun_out/r2/sweep_sw1_pair.txt | cut -c1-250
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2/bench5.json 2> gpurun_out/r2/bench5.err
python -c "
import json
d=json.load(open('gpurun_out/r2/bench5.json'))
print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'sampling', d['sampling']['value'], 'kernel us', d['roofline']['us_per_launch'], d['roofline']['frac'])
"
tail -2 gpurun_out/r2/bench5.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --mode parity_tc > gpurun_out/r2/bench5_ptc.json 2> gpurun_out/r2/bench5_ptc.err
python -c "
import json
d=json.load(open('gpurun_out/r2/bench5_ptc.json'))
print('parity_tc ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'sampling', d['sampling']['value'])
"
tail -2 gpurun_out/r2/bench5_ptc.err
