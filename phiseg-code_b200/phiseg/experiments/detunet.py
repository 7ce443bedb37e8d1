"""detunet (reference: phiseg/experiments/detunet.py): deterministic U-Net baseline.  The selectors exist so the file
loads; phiseg_model.phiseg raises NotImplementedError for it (outside the hot-path scope)."""
import tensorflow as tf
from phiseg.experiments._base import configure
from phiseg.model_zoo import likelihoods, posteriors, priors
from tfwrapper import normalisation as tfnorm

globals().update(configure('detunet', log_dir_name='lidc2', latent_levels=1, zdim0=6, annotator_range=[0],
                           KL_divergence_loss_weight=None))

posterior = posteriors.dummy
likelihood = likelihoods.det_unet2D
prior = priors.dummy
layer_norm = tfnorm.batch_norm
optimizer = tf.train.AdamOptimizer
