"""Experiment modules (config-as-python-module, the reference's config API: phiseg/experiments/*.py) and the loader
that makes *unmodified reference experiment files* importable without TensorFlow."""
import importlib.util
import os
import sys


def load_experiment(path):
    """Counterpart of phiseg_train.py:39-42 (SourceFileLoader(...).load_module()).

    A reference experiment file does `from phiseg.model_zoo import ...`, `import tensorflow as tf` and
    `from tfwrapper import normalisation as tfnorm`.  While the file executes, those module names are pointed at
    this package's selector modules (and `tensorflow` at tf_compat unless a real TensorFlow is importable), so both
    the shipped experiments and files written for the reference load unchanged."""
    pkg = sys.modules[__name__.rsplit('.', 2)[0]]            # phiseg_code_b200
    base = pkg.__name__
    alias = {
        'phiseg': base + '.phiseg',
        'phiseg.model_zoo': base + '.phiseg.model_zoo',
        'phiseg.model_zoo.posteriors': base + '.phiseg.model_zoo.posteriors',
        'phiseg.model_zoo.priors': base + '.phiseg.model_zoo.priors',
        'phiseg.model_zoo.likelihoods': base + '.phiseg.model_zoo.likelihoods',
        'phiseg.experiments': base + '.phiseg.experiments',
        'phiseg.experiments._base': base + '.phiseg.experiments._base',
        'tfwrapper': base + '.tfwrapper',
        'tfwrapper.normalisation': base + '.tfwrapper.normalisation',
        'tensorflow': base + '.tf_compat',
    }
    saved = {}
    try:
        for k, target in alias.items():
            saved[k] = sys.modules.get(k)
            sys.modules[k] = importlib.import_module(target)
        name = 'phiseg_experiment_' + os.path.splitext(os.path.basename(path))[0]
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def experiment_path(name):
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), name + '.py')
