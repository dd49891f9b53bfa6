"""lirec_b200 — B200-native (sm_100a) implementation of the LIReC model hot path.

The compute lives in liblirec_b200.so (hand-written CUDA behind a C ABI, see
include/lirec_b200.h); this package is the host-side mirror of the reference's Python
surface (utils.arg_pars, mlp.model / train / test, mixed_utils, resume).
"""
__version__ = "0.1.0"
