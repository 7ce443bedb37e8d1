"""phiseg_7_1_1annot (reference: phiseg/experiments/phiseg_7_1_1annot.py)."""
from phiseg.experiments._base import configure

globals().update(configure('phiseg_7_1_1annot', latent_levels=1, annotator_range=[0]))
