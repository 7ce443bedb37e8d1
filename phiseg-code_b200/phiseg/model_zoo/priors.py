"""Architecture selectors with the reference's names (phiseg/model_zoo/priors.py).  An experiment file assigns one of
these to `prior`; the topology itself is laid down by engine.build_program."""


class _Arch:
    def __init__(self, arch):
        self.arch = arch
        self.__name__ = arch

    def __repr__(self):
        return '<priors.%s>' % self.arch


phiseg = _Arch('phiseg')            # priors.py: hierarchical, one latent per resolution level
prob_unet2D = _Arch('probunet')     # priors.py: Probabilistic U-Net (Kohl et al.)
dummy = _Arch('dummy')                # priors.py: placeholder used by detunet
