"""Shared LIDC settings of the shipped experiments (values of phiseg/experiments/phiseg_7_5.py:5-56).
Each experiment module calls configure(...) with what it overrides and publishes the result as module attributes,
which is the config API phiseg_model.phiseg reads."""


def _selectors(nets, norm, optimizer):
    """The symbols an experiment assigns to posterior / prior / likelihood / layer_norm / optimizer."""
    from ..model_zoo import likelihoods, posteriors, priors
    from ...tfwrapper import normalisation as tfnorm
    from ... import tf_compat
    net = {'phiseg': (posteriors.phiseg, priors.phiseg, likelihoods.phiseg),
           'prob_unet2D': (posteriors.prob_unet2D, priors.prob_unet2D, likelihoods.prob_unet2D),
           'det_unet2D': (posteriors.dummy, priors.dummy, likelihoods.det_unet2D)}[nets]
    return dict(posterior=net[0], prior=net[1], likelihood=net[2],
                layer_norm={'batch_norm': tfnorm.batch_norm, 'group_norm2D': tfnorm.group_norm2D}[norm],
                optimizer={'adam': tf_compat.train.AdamOptimizer, 'momentum': tf_compat.train.MomentumOptimizer}[optimizer])


def configure(experiment_name, nets='phiseg', norm='batch_norm', optimizer='adam', **over):
    nlabels = over.get('nlabels', 2)
    num_labels_per_subject = over.get('num_labels_per_subject', 4)
    cfg = dict(
        experiment_name=experiment_name,
        log_dir_name='lidc',
        use_logistic_transform=False,
        latent_levels=5, resolution_levels=7, n0=32, zdim0=2, max_channel_power=4,
        data_identifier='lidc',
        preproc_folder='/srv/glusterfs/baumgach/preproc_data/lidc',
        data_root='/itet-stor/baumgach/bmicdatasets-originals/Originals/LIDC-IDRI/data_lidc.pickle',
        dimensionality_mode='2D',
        image_size=(128, 128, 1),
        nlabels=nlabels,
        num_labels_per_subject=num_labels_per_subject,
        augmentation_options={'do_flip_lr': True, 'do_flip_ud': True, 'do_rotations': True, 'do_scaleaug': True,
                              'nlabels': nlabels},
        lr_schedule_dict={0: 1e-3},
        deep_supervision=True,
        batch_size=12,
        num_iter=5000000,
        annotator_range=range(num_labels_per_subject),
        KL_divergence_loss_weight=1.0,
        exponential_weighting=True,
        residual_multinoulli_loss_weight=1.0,
        do_image_summaries=True,
        rescale_RGB=False,
        validation_frequency=500,
        validation_samples=16,
        num_validation_images=100,
        tensorboard_update_frequency=100,
    )
    cfg.update(_selectors(nets, norm, optimizer))
    cfg.update(over)
    return cfg
