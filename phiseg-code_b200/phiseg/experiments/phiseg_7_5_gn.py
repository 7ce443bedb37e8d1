"""phiseg_7_5 with group normalisation (tfwrapper/normalisation.py:17-36): per-sample statistics, so data-parallel
training is exactly equivalent to one large batch.  Not shipped by the reference; named by BASELINE.json north_star."""
import tensorflow as tf
from phiseg.experiments._base import configure
from phiseg.model_zoo import likelihoods, posteriors, priors
from tfwrapper import normalisation as tfnorm

globals().update(configure('phiseg_7_5_gn'))

posterior = posteriors.phiseg
likelihood = likelihoods.phiseg
prior = priors.phiseg
layer_norm = tfnorm.group_norm2D
optimizer = tf.train.AdamOptimizer
