"""phiseg/model_zoo/priors.py selectors (see model_zoo/__init__.py)."""
from . import Arch

phiseg = Arch('priors', 'phiseg', 'phiseg')                  # priors.py:51-128
prob_unet2D = Arch('priors', 'prob_unet2D', 'probunet')      # priors.py:8-48
dummy = Arch('priors', 'dummy', 'dummy')                     # priors.py:130-133
