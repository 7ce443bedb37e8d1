"""Architecture selectors with the reference's names (phiseg/model_zoo/posteriors.py).  An experiment file assigns one of
these to `posterior`; the topology itself is laid down by engine.build_program."""


class _Arch:
    def __init__(self, arch):
        self.arch = arch
        self.__name__ = arch

    def __repr__(self):
        return '<posteriors.%s>' % self.arch


phiseg = _Arch('phiseg')            # posteriors.py: hierarchical, one latent per resolution level
prob_unet2D = _Arch('probunet')     # posteriors.py: Probabilistic U-Net (Kohl et al.)
dummy = _Arch('dummy')                # posteriors.py: placeholder used by detunet
