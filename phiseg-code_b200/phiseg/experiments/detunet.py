"""detunet (reference: phiseg/experiments/detunet.py): deterministic U-Net baseline.  The selectors exist so the file
loads; phiseg_model.phiseg raises NotImplementedError for it (outside the hot-path scope)."""
from phiseg.experiments._base import configure

globals().update(configure('detunet', nets='det_unet2D', log_dir_name='lidc2', latent_levels=1, zdim0=6, annotator_range=[0],
                           KL_divergence_loss_weight=None))
