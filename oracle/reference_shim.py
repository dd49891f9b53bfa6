"""Import the UNMODIFIED reference (/root/reference) as a live oracle, in this container only.

TEST INFRASTRUCTURE — see oracle/__init__.py.  The reference tree is read-only and does not
travel to the GPU box; everything here degrades to `available() == False` when it is absent.

Shims:
  * utils/arg_pars.py parses sys.argv at import -> argv is cleared around the import;
  * mixed_utils/classification_dataloader.py imports plotly and (through text_utils)
    pytorch_pretrained_bert, neither installed nor used on this path -> stub modules;
  * torch >= 1.2 made `~uint8` bitwise and rejects uint8 mask indexing, which breaks
    MarginLoss / MarginTrackRelsLoss (mlp/model.py:459-460, 510-524); torch.ByteTensor and
    Tensor.byte are mapped to bool inside `torch11_masks()` to restore torch-1.1 semantics.
"""
import contextlib
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("LIREC_REFERENCE_ROOT", "/root/reference")

_state = {"opt": None, "model": None}


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "mlp", "model.py"))


@contextlib.contextmanager
def torch11_masks():
    old_bt, old_byte = torch.ByteTensor, torch.Tensor.byte
    torch.ByteTensor = lambda a: torch.as_tensor(a).bool()
    torch.Tensor.byte = torch.Tensor.bool
    try:
        yield
    finally:
        torch.ByteTensor, torch.Tensor.byte = old_bt, old_byte


def load():
    """Returns (opt, mlp.model module) of the reference; imports them on first use."""
    if _state["model"] is not None:
        return _state["opt"], _state["model"]
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    for name in ("plotly", "plotly.graph_objs", "plotly.graph_objs.layout"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["plotly.graph_objs.layout"].scene = None
    if "pytorch_pretrained_bert" not in sys.modules:
        m = types.ModuleType("pytorch_pretrained_bert")
        m.BertTokenizer = m.BertModel = m.BertForMaskedLM = None
        sys.modules["pytorch_pretrained_bert"] = m
    # the repo's own packages are named like the reference's (utils, mlp, ...): make sure the
    # reference's win for this import and are kept under their own module objects
    saved_argv, saved_path = sys.argv, list(sys.path)
    shadow = {k: sys.modules.pop(k) for k in list(sys.modules)
              if k.split(".")[0] in ("utils", "mlp", "mixed_utils", "text_utils", "visual_utils", "resume")}
    sys.argv = ["oracle"]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        from utils.arg_pars import opt  # noqa
        opt.device = "cpu"
        opt.text_dim, opt.visual_dim, opt.track_dim = 768, 2048, 2048
        opt.tracks = True
        opt.mlp_dim = 768 + 2048 + 2 * 2048
        import mlp.model as ref_model  # noqa
        _state["opt"], _state["model"] = opt, ref_model
        _state["modules"] = {k: v for k, v in sys.modules.items()
                             if k.split(".")[0] in ("utils", "mlp", "mixed_utils", "text_utils", "visual_utils")}
    finally:
        sys.argv, sys.path[:] = saved_argv, saved_path
        for k in list(sys.modules):
            if k.split(".")[0] in ("utils", "mlp", "mixed_utils", "text_utils", "visual_utils", "resume"):
                del sys.modules[k]
        sys.modules.update(shadow)
    return _state["opt"], _state["model"]


PRESETS = {
    # flag presets of resume/modalties.py:79-100, int_rels.py:88-115, int_ch.py:77-117,
    # int_rel_ch.py:87-124
    "modalities": dict(mod_check=True, tr_maximize=False, ints=1, ctx=0, gates=0, rels_multitask=False,
                       rels_multi_clip=False, modality="m", tracks=True),
    "int_rels": dict(mod_check=False, tr_maximize=False, ints=1, ctx=1, gates=1, rels_multitask=True,
                     rels_multi_clip=True, rels_n_clips=18, lymbda=1, tracks=True),
    "int_ch": dict(mod_check=False, tr_maximize=True, ints=1, ctx=0, gates=0, rels_multitask=False,
                   rels_multi_clip=False, tracks=True),
    "int_rel_ch": dict(mod_check=False, tr_maximize=True, ints=1, ctx=1, gates=1, rels_multitask=True,
                       rels_multi_clip=True, rels_n_clips=18, tracks=True),
}


def set_preset(name, **overrides):
    opt, _ = load()
    base = dict(tr_correct=False, tr_max_neg=False, tr_cat_distr=False, tr_sum_max_flag=True, dropout=0.3,
                margin=0.101, tr_margin=0.101, lymbda=1, modality="m", device="cpu")
    base.update(PRESETS[name])
    base.update(overrides)
    for k, v in base.items():
        setattr(opt, k, v)
    return opt


def create_model(name, n_classes, n_rels, seed=0, **overrides):
    """(model, loss) of the reference for a preset, random init under `seed`, printing silenced."""
    import io
    opt, ref_model = load()
    set_preset(name, **overrides)
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        model, loss, _ = ref_model.create_model(n_classes, n_rels=n_rels)
    return model, loss


def run_loss(loss, output, batch):
    """Call a reference loss under torch-1.1 mask semantics."""
    with torch11_masks():
        return loss(output, batch)
