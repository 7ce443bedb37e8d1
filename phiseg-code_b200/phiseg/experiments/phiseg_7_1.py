"""phiseg_7_1 (reference: phiseg/experiments/phiseg_7_1.py)."""
from phiseg.experiments._base import configure

globals().update(configure('phiseg_7_1', latent_levels=1))
