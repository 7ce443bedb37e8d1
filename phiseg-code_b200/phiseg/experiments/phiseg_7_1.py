"""phiseg_7_1 (reference: phiseg/experiments/phiseg_7_1.py)."""
import tensorflow as tf
from phiseg.experiments._base import configure
from phiseg.model_zoo import likelihoods, posteriors, priors
from tfwrapper import normalisation as tfnorm

globals().update(configure('phiseg_7_1', latent_levels=1))

posterior = posteriors.phiseg
likelihood = likelihoods.phiseg
prior = priors.phiseg
layer_norm = tfnorm.batch_norm
optimizer = tf.train.AdamOptimizer
