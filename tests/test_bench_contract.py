"""CPU check of the bench.py contract for the reference arm (`--impl reference`: the PyTorch-CPU oracle restatement of the
TF-1.12 path timed on the host cores): one JSON line with the keys the driver reads, and ranks other than 0 print
nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                           '--warmup', '1'], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    out = _run()
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'images/s' and d['higher_is_better'] is True
    assert d['metric'] == 'LIDC 128x128 training images/sec' and d['value'] > 0 and d['ms_per_step'] > 0
    assert d['vs_baseline'] is None and d['data'] == 'synthetic' and 'workload' in d['config']
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and 'batch 12' in cb['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_reference_arm_other_ranks_stay_silent():
    out = _run({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert out.returncode == 0 and out.stdout.strip() == ''


def test_supervisor_retries_a_stalled_run_in_the_conservative_configuration():
    """bench.py (one GPU) runs the measurement in a child process: a child that stalls is killed and the run repeated once
    with programmatic dependent launch and CTA pairs switched off; the line that comes out names the configuration that
    produced it.  Exercised with the built-in self-test hook (the default configuration sleeps, the conservative one
    answers) - no GPU involved."""
    env = dict(os.environ, BENCH_SELFTEST='stall-then-ok', BENCH_ATTEMPT_TIMEOUT='20')
    env.pop('BENCH_CHILD', None)
    env.pop('PHS_PDL', None)
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1', '--warmup', '1'],
                         capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['value'] == 1.0 and 'PHS_PDL=0' in d['config']['launch_config'] and 'PHS_HALO_PAIR=0' in d['config']['launch_config']
    assert 'did not finish' in out.stderr


def test_supervisor_reports_failure_when_both_attempts_fail():
    """Without a CUDA device the product path fails loudly (no CPU fallback): both attempts fail, nothing is printed on
    stdout, the exit code is not zero."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip('needs a machine without a GPU')
    env = dict(os.environ, BENCH_ATTEMPT_TIMEOUT='300')
    env.pop('BENCH_CHILD', None)
    env.pop('BENCH_SELFTEST', None)
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1', '--warmup', '1', '--no-cpu'],
                         capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert out.returncode != 0 and not [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert out.stderr.count('failed (exit code') == 2
