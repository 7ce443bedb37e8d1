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


def test_every_programmatically_launched_kernel_waits_for_its_predecessor():
    """Kernels launched through phs_launch carry the programmatic-stream-serialization attribute: they may be scheduled
    while the previous kernel of the stream still runs, so each of them MUST execute PHS_PDL_PROLOGUE() (griddepcontrol.wait)
    before touching global memory.  Source check: every kernel name handed to phs_launch has the macro in its body, before
    any __ldg / global pointer dereference that could matter (the macro is required to be among the first statements or
    right after the shared-memory / tensor-memory set-up)."""
    import re
    csrc = os.path.join(ROOT, 'phiseg-code_b200', 'csrc')
    src = {f: open(os.path.join(csrc, f)).read() for f in os.listdir(csrc) if f.endswith(('.cu', '.cuh'))}
    launched = set()
    for txt in src.values():
        for m in re.finditer(r'phs_launch(?:_tc|_cluster2)?\(\s*([A-Za-z_]\w*)', txt):
            launched.add(m.group(1))
    launched -= {'kernel', 'void'}      # the helper's own definition in common.cuh
    assert len(launched) >= 20
    for k in sorted(launched):
        body = None
        for txt in src.values():
            m = re.search(r'__global__[^;{]*?\b' + k + r'\s*\(', txt)
            if m:
                i = txt.index('{', m.end())
                depth, j = 1, i + 1
                while depth:
                    depth += {'{': 1, '}': -1}.get(txt[j], 0)
                    j += 1
                body = txt[i:j]
                break
        assert body is not None, 'kernel %s not found' % k
        if 'tmem_alloc' in body:
            # tensor-memory kernels: wait in the prologue, but release the successor only once the CTA HOLDS its tensor memory
            # (behind the barrier that follows the allocation) - a trigger in front of a blocking tcgen05.alloc lets the
            # successor's early CTAs take the columns this kernel is waiting for (common.cuh, PHS_PDL_WAIT)
            assert 'PHS_PDL_PROLOGUE()' not in body, k
            w, a, t = body.index('PHS_PDL_WAIT()'), body.index('tmem_alloc'), body.index('PHS_PDL_TRIGGER()')
            sync = min(i for i in (body.find('__syncthreads()', a), body.find('cluster_sync_all()', a)) if i >= 0)
            assert a < sync < t and w < t, 'kernel %s triggers its successor before it holds tensor memory' % k
        else:
            assert 'PHS_PDL_PROLOGUE()' in body, 'kernel %s is launched with programmatic serialization but never waits' % k
