"""probunet (reference: phiseg/experiments/probunet.py)."""
import tensorflow as tf
from phiseg.experiments._base import configure
from phiseg.model_zoo import likelihoods, posteriors, priors
from tfwrapper import normalisation as tfnorm

globals().update(configure('probunet', latent_levels=1, zdim0=6))

posterior = posteriors.prob_unet2D
likelihood = likelihoods.prob_unet2D
prior = priors.prob_unet2D
layer_norm = tfnorm.batch_norm
optimizer = tf.train.AdamOptimizer
