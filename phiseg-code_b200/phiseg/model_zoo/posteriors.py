"""phiseg/model_zoo/posteriors.py selectors (see model_zoo/__init__.py)."""
from . import Arch

phiseg = Arch('posteriors', 'phiseg', 'phiseg')              # posteriors.py:56-132: hierarchical, one latent per level
prob_unet2D = Arch('posteriors', 'prob_unet2D', 'probunet')  # posteriors.py:9-53: Probabilistic U-Net (Kohl et al.)
dummy = Arch('posteriors', 'dummy', 'dummy')                 # posteriors.py:135-138: placeholder used by detunet
