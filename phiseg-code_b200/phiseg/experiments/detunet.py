"""detunet (reference: phiseg/experiments/detunet.py): deterministic U-Net baseline (likelihoods.det_unet2D with the dummy
posterior / prior: no latent variables, cross-entropy only)."""
from phiseg.experiments._base import configure

globals().update(configure('detunet', nets='det_unet2D', log_dir_name='lidc2', latent_levels=1, zdim0=6, annotator_range=[0],
                           KL_divergence_loss_weight=None))
