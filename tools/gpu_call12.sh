set -x
mkdir -p gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
PHS_DP_MODE=graph timeout 110 $TR --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 4 --no-cpu > gpurun_out/r2/bench_2gpu_graph.json 2> gpurun_out/r2/bench_2gpu_graph.err
echo rc=$?
grep -E "^\{" gpurun_out/r2/bench_2gpu_graph.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('2 GPUs graph: value', d['value'], 'ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
"
PHS_PDL=0 PHS_DP_MODE=graph timeout 110 $TR --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 4 --no-cpu > gpurun_out/r2/bench_2gpu_graph_nopdl.json 2> gpurun_out/r2/bench_2gpu_graph_nopdl.err
echo rc=$?
grep -E "^\{" gpurun_out/r2/bench_2gpu_graph_nopdl.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('2 GPUs graph no PDL: value', d['value'], 'ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
"
