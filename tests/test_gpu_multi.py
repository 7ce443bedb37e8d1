"""-m gpu, needs >= 2 GPUs (skipped on a single-GPU box; run with `gpurun --gpus 2`): the real engine under data
parallelism - bucketed NCCL all-reduce of the flat gradient buffer inside the captured step - against one GPU on the
concatenated batch."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize('mode,tol', [('parity', 2e-5), ('fast', 2e-2)])
def test_data_parallel_step_equals_single_gpu_large_batch(mode, tol):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()), os.path.join(ROOT, 'tests', 'dp_worker.py'), mode, str(tol)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    sys.stdout.write(out.stdout[-3000:])
    assert out.returncode == 0, out.stderr[-3000:]
