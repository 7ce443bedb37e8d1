set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29531 tests/dp_worker.py parity 2e-5 2>&1 | grep -E "rank|dp-equiv|Error|error|Traceback" | tail -6
timeout 200 $TR --master-port 29532 tests/dp_worker.py fast 2e-2 2>&1 | grep -E "rank|dp-equiv|Error|error|Traceback" | tail -6
PHS_DP_MODE=graph timeout 120 $TR --master-port 29533 tests/dp_worker.py parity 2e-5 2>&1 | grep -E "rank|dp-equiv|Error|error|Traceback" | tail -6
