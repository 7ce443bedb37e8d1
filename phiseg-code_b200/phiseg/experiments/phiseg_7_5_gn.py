"""phiseg_7_5 with group normalisation (tfwrapper/normalisation.py:17-36): per-sample statistics, so data-parallel
training is exactly equivalent to one large batch.  Not shipped by the reference; named by BASELINE.json north_star."""
from phiseg.experiments._base import configure

globals().update(configure('phiseg_7_5_gn', norm='group_norm2D'))
