"""Synthetic LIDC-shaped inputs (the dataset itself is not available offline; data/lidc_data_loader.py:92 feeds
image - 0.5 and uint8 annotation masks)."""
from .synthetic import SyntheticLIDC, synthetic_batch, synthetic_eps  # noqa: F401
