"""BASELINE.json config 5: phiseg_7_5 on synthetic 256x256 4-class images (deep-resolution stress)."""
import tensorflow as tf
from phiseg.experiments._base import configure
from phiseg.model_zoo import likelihoods, posteriors, priors
from tfwrapper import normalisation as tfnorm

globals().update(configure('phiseg_7_5_256', image_size=(256, 256, 1), nlabels=4, batch_size=32))

posterior = posteriors.phiseg
likelihood = likelihoods.phiseg
prior = priors.phiseg
layer_norm = tfnorm.batch_norm
optimizer = tf.train.AdamOptimizer
