"""CPU check of the multiply-high division constants behind the index arithmetic of the streaming kernels
(csrc/common.cuh: fdiv_make / idx4_make): a host program built with nvcc compares them with `/` for every divisor up to
4096 and 200k random (divisor, index) pairs.  No GPU needed."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fdiv_constants_match_integer_division(tmp_path):
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        pytest.skip('nvcc not available')
    exe = str(tmp_path / 'fdiv_check')
    src = os.path.join(ROOT, 'tests', 'native', 'fdiv_check.cu')
    subprocess.check_call([nvcc, '-O2', '-std=c++17', '-o', exe, src])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.startswith('OK'), out.stdout + out.stderr
