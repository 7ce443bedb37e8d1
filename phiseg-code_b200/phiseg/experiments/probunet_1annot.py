"""probunet_1annot (reference: phiseg/experiments/probunet_1annot.py)."""
import tensorflow as tf
from phiseg.experiments._base import configure
from phiseg.model_zoo import likelihoods, posteriors, priors
from tfwrapper import normalisation as tfnorm

globals().update(configure('probunet_1annot', latent_levels=1, zdim0=6, annotator_range=[0]))

posterior = posteriors.prob_unet2D
likelihood = likelihoods.prob_unet2D
prior = priors.prob_unet2D
layer_norm = tfnorm.batch_norm
optimizer = tf.train.AdamOptimizer
