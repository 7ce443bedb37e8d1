set -x
mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2/bench_2gpu.json 2> gpurun_out/r2/bench_2gpu.err
tail -5 gpurun_out/r2/bench_2gpu.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2/bench_2gpu.json') if l.startswith('{')][-1])
print('2 GPUs: value', d['value'], 'ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])
"
