"""Architecture selectors with the reference's names (phiseg/model_zoo/likelihoods.py).  An experiment file assigns one of
these to `likelihood`; the topology itself is laid down by engine.build_program."""


class _Arch:
    def __init__(self, arch):
        self.arch = arch
        self.__name__ = arch

    def __repr__(self):
        return '<likelihoods.%s>' % self.arch


phiseg = _Arch('phiseg')            # likelihoods.py: hierarchical, one latent per resolution level
prob_unet2D = _Arch('probunet')     # likelihoods.py: Probabilistic U-Net (Kohl et al.)
det_unet2D = _Arch('det_unet')      # likelihoods.py:10-79: deterministic U-Net (the prob. U-Net's U-Net without z)
