"""BASELINE.json config 5: phiseg_7_5 on synthetic 256x256 4-class images (deep-resolution stress)."""
from phiseg.experiments._base import configure

globals().update(configure('phiseg_7_5_256', image_size=(256, 256, 1), nlabels=4, batch_size=32))
