"""lirec_b200 — B200-native (sm_100a) implementation of the LIReC model hot path.

The compute lives in liblirec_b200.so (hand-written CUDA behind a C ABI, see
include/lirec_b200.h); this package is the host-side mirror of the reference's Python
surface (utils.arg_pars, mlp.model / train / test, mixed_utils, resume).
"""
__version__ = "0.1.0"


def install_aliases():
    """Expose the package's sub-packages under the reference's top-level module names (`utils`, `mlp`,
    `mixed_utils`, `resume`), so code written against the reference — `from utils.arg_pars import opt`,
    `import mlp.model` — runs on this implementation unchanged.  Same module objects: one `opt`."""
    import importlib
    import sys
    for name in ("utils", "utils.arg_pars", "utils.util_functions", "utils.model_saver", "mlp", "mlp.model",
                 "mlp.train", "mlp.test", "mixed_utils", "mixed_utils.update_arg_pars",
                 "mixed_utils.classification_dataloader", "mixed_utils.mixed_features", "resume",
                 "resume.modalties", "resume.int_rels", "resume.int_ch", "resume.int_rel_ch"):
        sys.modules[name] = importlib.import_module("lirec_b200." + name)
