"""Architecture selectors with the reference's names (phiseg/model_zoo/{posteriors,priors,likelihoods}.py).

An experiment file assigns one symbol of each module to `posterior` / `prior` / `likelihood` (phiseg/experiments/*.py); in
the reference those symbols are graph-building functions that phiseg_model.py:37-98 calls.  Here the topology is laid down
as a static launch program by engine.build_program, so the symbols only SELECT it: phiseg_model.net_config_from_experiment
reads `.arch` ('phiseg' | 'probunet' | 'det_unet' | 'dummy').  Calling one raises with that explanation instead of failing
obscurely."""


class Arch:
    def __init__(self, module, name, arch):
        self.arch = arch
        self.__name__ = name
        self._module = module

    def __repr__(self):
        return '<%s.%s>' % (self._module, self.__name__)

    def __call__(self, *args, **kwargs):
        raise TypeError('%r selects an architecture; the network is built by engine.build_program (use phiseg_model.phiseg '
                        'and its generate_* / predict methods to evaluate it)' % (self,))
