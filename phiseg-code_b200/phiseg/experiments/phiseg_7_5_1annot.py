"""phiseg_7_5_1annot (reference: phiseg/experiments/phiseg_7_5_1annot.py)."""
import tensorflow as tf
from phiseg.experiments._base import configure
from phiseg.model_zoo import likelihoods, posteriors, priors
from tfwrapper import normalisation as tfnorm

globals().update(configure('phiseg_7_5_1annot', annotator_range=[0]))

posterior = posteriors.phiseg
likelihood = likelihoods.phiseg
prior = priors.phiseg
layer_norm = tfnorm.batch_norm
optimizer = tf.train.AdamOptimizer
