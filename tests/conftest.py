import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# the reference-compatible flag module parses argv at import; keep pytest's argv away from it
sys.argv = sys.argv[:1]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu")


@pytest.fixture(scope="session")
def built_lib():
    """liblirec_b200.so, built in-tree (nvcc cross-compiles without a GPU)."""
    from lirec_b200 import build
    return build.build()


@pytest.fixture()
def opt_preset():
    """Set the reference flag presets (resume/*.py) on the global opt; restores on exit."""
    from lirec_b200.utils.arg_pars import opt
    saved = dict(vars(opt))
    presets = {
        "modalities": dict(mod_check=True, tr_maximize=False, ints=1, ctx=0, gates=0, rels_multitask=False),
        "int_rels": dict(mod_check=False, tr_maximize=False, ints=1, ctx=1, gates=1, rels_multitask=True,
                         rels_multi_clip=True, rels_n_clips=18),
        "int_ch": dict(mod_check=False, tr_maximize=True, ints=1, ctx=0, gates=0, rels_multitask=False),
        "int_rel_ch": dict(mod_check=False, tr_maximize=True, ints=1, ctx=1, gates=1, rels_multitask=True,
                           rels_multi_clip=True, rels_n_clips=18),
    }

    def apply(name, **over):
        base = dict(tracks=True, modality="m", device="cuda", tr_correct=False, tr_max_neg=False,
                    tr_cat_distr=False, tr_sum_max_flag=True, dropout=0.3, margin=0.101, tr_margin=0.101,
                    lymbda=1.0, fused_adam=0)
        base.update(presets[name])
        base.update(over)
        for k, v in base.items():
            setattr(opt, k, v)
        return opt

    yield apply
    for k in list(vars(opt)):
        if k not in saved:
            delattr(opt, k)
    for k, v in saved.items():
        setattr(opt, k, v)
