"""CPU-side checks of the drop-in boundary: the shared library builds (nvcc cross-compiles sm_100a without a GPU),
loads, and exports every function include/phiseg_sm100.h declares; the ctypes binding declares the same set; the
experiment modules expose the attributes the reference's phiseg class reads.  No compute calls."""
import importlib
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'phiseg_sm100.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(phs_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol(lib):
    h = lib.load()
    names = header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(h, n), 'libphiseg_sm100.so does not export %s' % n
    assert sorted(lib.exported_symbols()) == names, 'lib.py and include/phiseg_sm100.h disagree'
    assert h.phs_arch() == 100 and h.phs_version() >= 100


def test_missing_library_fails_loudly(lib, monkeypatch):
    monkeypatch.setattr(lib, '_lib', None)
    monkeypatch.setattr(lib, 'LIB_PATH', os.path.join(ROOT, 'does', 'not', 'exist.so'))
    with pytest.raises(ImportError):
        lib.load()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'phiseg-code_b200')
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                txt = open(os.path.join(d, f)).read()
                assert 'oracle' not in txt.replace('load_oracle', '').lower() or f == 'build.py', os.path.join(d, f)


@pytest.mark.parametrize('name', ['phiseg_7_5', 'phiseg_7_1', 'probunet', 'phiseg_7_5_gn', 'phiseg_7_5_256'])
def test_experiment_modules_expose_reference_attributes(pkg, name):
    ex = importlib.import_module('phiseg_code_b200.phiseg.experiments')
    exp = ex.load_experiment(ex.experiment_path(name))
    for attr in ('image_size', 'nlabels', 'zdim0', 'n0', 'resolution_levels', 'latent_levels', 'layer_norm', 'posterior',
                 'prior', 'likelihood', 'residual_multinoulli_loss_weight', 'KL_divergence_loss_weight',
                 'exponential_weighting', 'optimizer', 'lr_schedule_dict', 'batch_size', 'num_iter',
                 'validation_frequency', 'log_dir_name', 'experiment_name'):
        assert hasattr(exp, attr), (name, attr)
    eng = importlib.import_module('phiseg_code_b200.engine')
    pm_cfg = importlib.import_module('phiseg_code_b200.phiseg.phiseg_model').net_config_from_experiment(exp, 'fast')
    spec = eng.build_spec(pm_cfg)
    names = [n for n, _, _ in spec]
    assert len(names) == len(set(names))
    if pm_cfg.arch == 'phiseg':
        assert 'posterior/z0_pre_1/W' in names and 'likelihood/y_lvl0/b' in names
