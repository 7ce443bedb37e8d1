"""phiseg/model_zoo/likelihoods.py selectors (see model_zoo/__init__.py)."""
from . import Arch

phiseg = Arch('likelihoods', 'phiseg', 'phiseg')             # likelihoods.py:162-223
prob_unet2D = Arch('likelihoods', 'prob_unet2D', 'probunet') # likelihoods.py:81-159
det_unet2D = Arch('likelihoods', 'det_unet2D', 'det_unet')   # likelihoods.py:10-79: the prob. U-Net's U-Net without z
