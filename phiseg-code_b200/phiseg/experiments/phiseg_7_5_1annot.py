"""phiseg_7_5_1annot (reference: phiseg/experiments/phiseg_7_5_1annot.py)."""
from phiseg.experiments._base import configure

globals().update(configure('phiseg_7_5_1annot', annotator_range=[0]))
