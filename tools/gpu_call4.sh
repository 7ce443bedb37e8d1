set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "reproducible or fast_mode or full_size" 2>&1 | tail -5
python tools/bench_conv.py > gpurun_out/r2/bench_conv2.txt 2>&1; cat gpurun_out/r2/bench_conv2.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2/bench2.json 2> gpurun_out/r2/bench2.err
python -c "
import json
d=json.load(open('gpurun_out/r2/bench2.json'))
print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'sampling', d['sampling']['value'], 'kernel us', d['roofline']['us_per_launch'], d['roofline']['frac'])
"
tail -3 gpurun_out/r2/bench2.err
