"""Drop-in check of the config API against the reference's own experiment files: the UNMODIFIED
files under /root/reference/phiseg/experiments are loaded through experiments.load_experiment (TensorFlow and the
reference packages resolve to this package's selector modules) and compared attribute by attribute with the shipped
modules of the same name, and with the network configuration the model derives from them.  Runs only where the reference
checkout exists (this container); it is not needed, and skipped, on the GPU box."""
import importlib
import os

import pytest

REF = '/root/reference/phiseg/experiments'
NAMES = ['phiseg_7_5', 'phiseg_7_1', 'probunet', 'phiseg_7_5_1annot', 'phiseg_7_1_1annot', 'probunet_1annot', 'detunet']


def _public(mod):
    out = {}
    for k, v in vars(mod).items():
        if k.startswith('_') or isinstance(v, type(os)) or k in ('configure',):
            continue
        out[k] = v
    return out


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference checkout not present')
@pytest.mark.parametrize('name', NAMES)
def test_unmodified_reference_experiment_loads_and_matches(pkg, name):
    ex = importlib.import_module('phiseg_code_b200.phiseg.experiments')
    pm = importlib.import_module('phiseg_code_b200.phiseg.phiseg_model')
    ref = ex.load_experiment(os.path.join(REF, name + '.py'))
    mine = ex.load_experiment(ex.experiment_path(name))
    a, b = _public(ref), _public(mine)
    # every attribute the reference file defines exists here with the same value (callables: the same selector symbol)
    for k, v in a.items():
        assert k in b, '%s: attribute %s of the reference experiment is missing' % (name, k)
        w = b[k]
        if callable(v) or callable(w):
            assert getattr(v, '__name__', v) == getattr(w, '__name__', w), (name, k, v, w)
        elif isinstance(v, range):
            assert list(v) == list(w), (name, k)
        else:
            assert v == w, (name, k, v, w)
    # and the engine derives the same network from both
    ca, cb = pm.net_config_from_experiment(ref, 'fast'), pm.net_config_from_experiment(mine, 'fast')
    for f in ('arch', 'H', 'W', 'Cx', 'nlabels', 'zdim0', 'n0', 'R', 'L', 'norm', 'KL_weight', 'xent_weight',
              'exponential_weighting', 'weight_decay', 'optimizer'):
        assert getattr(ca, f) == getattr(cb, f), (name, f, getattr(ca, f), getattr(cb, f))


@pytest.mark.skipif(not os.path.isfile('/root/reference/phiseg/phiseg_model.py'), reason='reference checkout not present')
def test_class_surface_covers_the_reference(pkg):
    """Every public method of the reference's phiseg class (phiseg/phiseg_model.py, read with ast: TensorFlow is not
    importable) exists on the replacement with the same leading argument names; extra keyword arguments are allowed."""
    import ast
    import inspect
    pm = importlib.import_module('phiseg_code_b200.phiseg.phiseg_model')
    tree = ast.parse(open('/root/reference/phiseg/phiseg_model.py').read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'phiseg'][0]
    # loss-graph builders are internal to the TF graph construction (their arithmetic lives in the kernels here)
    internal = {'KL_two_gauss_with_diag_cov', 'multinoulli_loss_with_logits', 'add_residual_multinoulli_loss',
                'add_hierarchical_KL_div_loss', 'add_weight_decay'}
    missing, mismatched = [], []
    for fn in cls.body:
        if not isinstance(fn, ast.FunctionDef) or fn.name.startswith('_') or fn.name in internal:
            continue
        if not hasattr(pm.phiseg, fn.name):
            missing.append(fn.name)
            continue
        ref_args = [a.arg for a in fn.args.args][1:]
        mine = [p for p in inspect.signature(getattr(pm.phiseg, fn.name)).parameters][1:]
        if mine[:len(ref_args)] != ref_args:
            mismatched.append((fn.name, ref_args, mine))
    assert not missing, 'reference methods without a counterpart: %s' % missing
    assert not mismatched, 'argument names differ: %s' % mismatched


@pytest.mark.skipif(not os.path.isfile('/root/reference/tfwrapper/layers.py'), reason='reference checkout not present')
def test_operator_layer_signatures_match_the_reference(pkg):
    """The op-level drop-in keeps the reference's argument names, order and defaults (read with ast)."""
    import ast
    import inspect

    def ref_sigs(path):
        tree = ast.parse(open(path).read())
        out = {}
        for fn in tree.body:
            if isinstance(fn, ast.FunctionDef):
                args = [a.arg for a in fn.args.args]
                defaults = [ast.literal_eval(d) if isinstance(d, (ast.Constant, ast.Tuple, ast.UnaryOp)) else '<expr>'
                            for d in fn.args.defaults]
                out[fn.name] = (args, defaults, fn.args.kwarg.arg if fn.args.kwarg else None)
        return out

    layers = importlib.import_module('phiseg_code_b200.tfwrapper.layers')
    utils = importlib.import_module('phiseg_code_b200.tfwrapper.utils')
    checks = [(ref_sigs('/root/reference/tfwrapper/layers.py'), layers,
               ['conv2D', 'averagepool2D', 'global_averagepool2D', 'bilinear_upsample2D', 'crop_and_concat']),
              (ref_sigs('/root/reference/tfwrapper/utils.py'), utils, ['get_weight_variable', 'get_bias_variable'])]
    for ref, mod, names in checks:
        for n in names:
            args, defaults, kwarg = ref[n]
            sig = inspect.signature(getattr(mod, n))
            mine = [p for p in sig.parameters.values() if p.kind == p.POSITIONAL_OR_KEYWORD]
            assert [p.name for p in mine] == args, (n, args, [p.name for p in mine])
            mine_def = [p.default for p in mine if p.default is not inspect.Parameter.empty]
            assert len(mine_def) == len(defaults), (n, defaults, mine_def)
            for d_ref, d_mine in zip(defaults, mine_def):
                if d_ref != '<expr>':
                    assert d_ref == d_mine, (n, d_ref, d_mine)
            has_kw = any(p.kind == p.VAR_KEYWORD for p in sig.parameters.values())
            assert has_kw == (kwarg is not None), (n, kwarg)


@pytest.mark.skipif(not os.path.isdir('/root/reference/phiseg/model_zoo'), reason='reference checkout not present')
def test_variable_names_follow_the_reference_scopes(pkg):
    """Every convolution scope name in the flat parameter buffer (engine.build_spec) matches a name / name pattern the
    reference's model_zoo passes to layers.conv2D or tf.variable_scope (string literals and '...%d...' % (...) format
    strings, read with ast), under the top-level scopes posterior / prior / likelihood."""
    import ast
    import re
    eng = importlib.import_module('phiseg_code_b200.engine')
    patterns = set()
    for f in ('posteriors.py', 'priors.py', 'likelihoods.py'):
        tree = ast.parse(open(os.path.join('/root/reference/phiseg/model_zoo', f)).read())
        for node in ast.walk(tree):
            s = None
            if isinstance(node, ast.BinOp) and isinstance(node.op, ast.Mod) and isinstance(node.left, ast.Constant) \
                    and isinstance(node.left.value, str):
                s = node.left.value
            elif isinstance(node, ast.Constant) and isinstance(node.value, str):
                s = node.value
            if s and re.fullmatch(r'[A-Za-z0-9_%]+', s):
                patterns.add(s)
    regs = [re.compile('^' + re.escape(p).replace('%d', r'\d+') + '$') for p in patterns]

    def known(part):
        return any(r.match(part) for r in regs)

    norm_tail = {'batch_norm', 'BatchNorm', 'group_norm', 'W', 'b', 'beta', 'gamma', 'moving_mean', 'moving_variance'}
    for arch, kw in (('phiseg', {}), ('probunet', dict(zdim0=6, latent_levels=1)),
                     ('det_unet', dict(zdim0=6, latent_levels=1, KL_weight=None))):
        cfg = eng.NetConfig(arch=arch, image_size=(128, 128, 1), mode='parity', norm='batch_norm', **kw)
        for name, shape, kind in eng.build_spec(cfg):
            parts = name.split('/')
            assert parts[0] in ('posterior', 'prior', 'likelihood'), name
            for part in parts[1:]:
                assert part in norm_tail or known(part), 'scope %r of %s has no counterpart in the reference' % (part, name)
