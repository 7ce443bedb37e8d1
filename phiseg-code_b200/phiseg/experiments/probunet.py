"""probunet (reference: phiseg/experiments/probunet.py)."""
from phiseg.experiments._base import configure

globals().update(configure('probunet', nets='prob_unet2D', latent_levels=1, zdim0=6))
