import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with `-m gpu`)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device visible in this container')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def pkg():
    from __graft_entry__ import load_package
    return load_package()


@pytest.fixture(scope='session')
def oracle():
    from __graft_entry__ import load_oracle
    return load_oracle()


@pytest.fixture(scope='session')
def lib(pkg):
    """The loaded C-ABI library (built in-tree by __graft_entry__.build())."""
    import importlib
    from __graft_entry__ import build
    if not os.path.exists(os.path.join(ROOT, 'phiseg-code_b200', 'libphiseg_sm100.so')):
        build()
    return importlib.import_module('phiseg_code_b200.lib')
