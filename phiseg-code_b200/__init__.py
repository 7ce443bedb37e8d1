"""phiseg-code_b200: B200-native (sm_100a) engine for the PHiSeg training / sampling hot path.

The directory name contains a hyphen, so the package is imported under the name ``phiseg_code_b200`` through
``__graft_entry__.load_package()`` (or any importlib spec pointing at this __init__).  Layout:

  csrc/                hand-written CUDA kernels + the C-ABI (include/phiseg_sm100.h)
  lib.py               ctypes binding of libphiseg_sm100.so (no CPU fallback)
  engine.py            static launch programs: forward, hand-written backward, optimizer
  phiseg/              host-side mirror of the reference's phiseg package (phiseg_model.phiseg, experiments, model_zoo)
  tfwrapper/           selectable symbols of the reference's tfwrapper (normalisation.batch_norm / group_norm2D, losses)
  tf_compat.py         the tf.train.* names experiment files reference
"""
__version__ = '0.1.0'
