"""phiseg_7_5 (reference: phiseg/experiments/phiseg_7_5.py)."""
from phiseg.experiments._base import configure

globals().update(configure('phiseg_7_5'))
