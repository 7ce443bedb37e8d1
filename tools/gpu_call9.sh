set -x
mkdir -p gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
# 1. default data-parallel path (grad graph -> all-reduce -> optimizer graph)
timeout 240 $TR --master-port 29521 tests/dp_worker.py parity 2e-5 2>&1 | grep -vE "^W1017|OMP_NUM" | tail -6
timeout 240 $TR --master-port 29522 bench.py --gpus 2 --steps 10 --warmup 4 --no-cpu > gpurun_out/r2/bench_2gpu.json 2> gpurun_out/r2/bench_2gpu.err
tail -2 gpurun_out/r2/bench_2gpu.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2/bench_2gpu.json') if l.startswith('{')][-1])
print('2 GPUs eager: value', d['value'], 'ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
"
# 2. can NCCL be captured here at all?
timeout 90 $TR --master-port 29523 tools/nccl_graph_probe.py main 2>&1 | grep -E "rank|Error|error" | tail -8
timeout 90 $TR --master-port 29524 tools/nccl_graph_probe.py side 2>&1 | grep -E "rank|Error|error" | tail -8
# 3. the captured, bucketed all-reduce of the engine
PHS_DP_MODE=graph timeout 150 $TR --master-port 29525 tests/dp_worker.py parity 2e-5 2>&1 | grep -vE "^W1017|OMP_NUM" | tail -6
