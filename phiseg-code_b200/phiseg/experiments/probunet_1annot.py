"""probunet_1annot (reference: phiseg/experiments/probunet_1annot.py)."""
from phiseg.experiments._base import configure

globals().update(configure('probunet_1annot', nets='prob_unet2D', latent_levels=1, zdim0=6, annotator_range=[0]))
