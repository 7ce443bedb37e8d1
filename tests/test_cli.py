"""The thin training CLI (phiseg_train.py at the repo root; reference phiseg_train.py:16-50): argument surface, experiment
loading by path + copy into the log directory, array loading - and, on the GPU, a few training iterations with validation
and checkpoints on synthetic LIDC-shaped data."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import phiseg_train  # noqa: E402

EXP = os.path.join(ROOT, 'phiseg-code_b200', 'phiseg', 'experiments', 'phiseg_7_5.py')


def test_cli_arguments_follow_the_reference():
    a = phiseg_train.parse_args([EXP])
    assert a.EXP_PATH == EXP and a.data is None and a.synthetic == 0 and a.num_iter is None and a.mode is None
    a = phiseg_train.parse_args([EXP, '--synthetic', '32', '--num-iter', '5', '--log-root', '/tmp/x', '--mode', 'parity'])
    assert (a.synthetic, a.num_iter, a.log_root, a.mode) == (32, 5, '/tmp/x', 'parity')
    with pytest.raises(SystemExit):
        phiseg_train.parse_args([])                 # EXP_PATH is positional and required, like the reference


def test_prepare_loads_the_experiment_and_copies_it(pkg, tmp_path):
    exp, log_dir = phiseg_train.prepare(EXP, str(tmp_path))
    assert exp.experiment_name == 'phiseg_7_5' and exp.image_size == (128, 128, 1) and exp.batch_size == 12
    assert log_dir == os.path.join(str(tmp_path), exp.log_dir_name, exp.experiment_name)
    assert os.path.exists(os.path.join(log_dir, 'phiseg_7_5.py'))          # phiseg_train.py:46
    assert exp.log_root == str(tmp_path)


@pytest.mark.skipif(not os.path.exists('/root/reference/phiseg/experiments/probunet.py'), reason='reference tree not present')
def test_prepare_takes_an_unmodified_reference_experiment(pkg, tmp_path):
    exp, log_dir = phiseg_train.prepare('/root/reference/phiseg/experiments/probunet.py', str(tmp_path))
    assert exp.experiment_name == 'probunet' and exp.latent_levels == 1 and exp.zdim0 == 6
    assert os.path.exists(os.path.join(log_dir, 'probunet.py'))


def test_load_arrays_npz(tmp_path):
    p = str(tmp_path / 'd.npz')
    rng = np.random.default_rng(0)
    np.savez(p, train_images=rng.random((6, 16, 16)), train_labels=rng.integers(0, 2, (6, 16, 16, 4), dtype=np.uint8),
             val_images=rng.random((2, 16, 16)), val_labels=rng.integers(0, 2, (2, 16, 16, 4), dtype=np.uint8))
    d = phiseg_train.load_arrays(p)
    assert set(d) == {'train', 'val'} and d['train']['images'].shape == (6, 16, 16) and d['val']['labels'].shape == (2, 16, 16, 4)
    np.savez(p, train_images=rng.random((6, 16, 16)), train_labels=rng.integers(0, 2, (6, 16, 16, 4), dtype=np.uint8))
    with pytest.raises(RuntimeError):
        phiseg_train.load_arrays(p)


@pytest.mark.gpu
def test_cli_trains_validates_and_checkpoints(pkg, tmp_path):
    exp, log_dir = phiseg_train.prepare(EXP, str(tmp_path))
    exp.image_size = (64, 64, 1)
    exp.batch_size = 4
    exp.num_iter = 5
    exp.validation_frequency = 2
    exp.num_validation_images = 2
    exp.validation_samples = 4
    model = phiseg_train.main(exp, synthetic=16)
    files = os.listdir(log_dir)
    assert any(f.startswith('model.ckpt-4') for f in files), files
    assert any(f.startswith('model_best_dice') for f in files) and any(f.startswith('model_best_ged') for f in files)
    assert np.isfinite(model.last_validation['ged']) and 0.0 <= model.last_validation['dice'] <= 1.0
    assert model.params.step == 5
