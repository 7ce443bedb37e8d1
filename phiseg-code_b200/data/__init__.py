"""Inputs of the training loop: the device-resident batch provider (data/batch_provider.py, data/lidc_data.py of the
reference) and synthetic LIDC-shaped data (the dataset itself is not available offline; data/lidc_data_loader.py:92
feeds image - 0.5 and uint8 annotation masks)."""
from .batch_provider import BatchProvider, lidc_data  # noqa: F401
from .synthetic import SyntheticLIDC, synthetic_batch, synthetic_eps  # noqa: F401
