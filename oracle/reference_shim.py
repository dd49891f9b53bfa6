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


class DropoutReplay(torch.nn.Module):
    """Stand-in for the reference's `self.dropout` modules (mlp/model.py:52, 347): instead of drawing
    from torch's global RNG it applies the NEXT mask of `queue` (0/1 tensors, any shape with the input's
    element count) scaled by 1/(1-p) — nn.Dropout's own train-mode arithmetic.  The queue is filled in the
    reference's call order (mlp/model.py:62-88, 155-196, 282-327, 353):
        ints: txt, vis, tracks1, tracks2, cat   ctx: txt, vis, tracks1, tracks2, cat   gate
    so the oracle's keyed masks and the unmodified reference forward see the same masks."""

    def __init__(self, p, queue):
        super().__init__()
        self.p, self.queue = p, queue

    def forward(self, x):
        m = self.queue.pop(0)
        assert m.numel() == x.numel(), (tuple(m.shape), tuple(x.shape))
        return x * m.reshape(x.shape).to(x.dtype) / (1.0 - self.p)


def mask_order(masks, modality="m", tracks=True):
    """Keys of oracle/model.py's `masks` dict in the order the reference's forward calls dropout."""
    slots = [s for s in ("txt", "vis", "tracks1", "tracks2")
             if (s == "txt" and modality in ("m", "t")) or (s == "vis" and modality in ("m", "v"))
             or (s.startswith("tracks") and tracks)]
    order = []
    for br in ("ints", "ctx"):
        if ("cat", br) not in masks:
            continue
        order += [("l1", br, s) for s in slots] + [("cat", br)]
    if ("gate",) in masks:
        order.append(("gate",))
    return order


def replay_dropout(model, masks, p, modality="m", tracks=True):
    """Put the reference model in train mode with its dropout modules replaced by DropoutReplay over
    `masks` (oracle key -> 0/1 tensor).  Returns the queue (empty after one forward)."""
    queue = [masks[k] for k in mask_order(masks, modality, tracks)]
    model.train()
    model.dropout = DropoutReplay(p, queue)
    if hasattr(model, "gates_ints"):
        model.gates_ints.dropout = DropoutReplay(p, queue)
    return queue


def random_masks(kind, rows, S, J, gate_dim, p, gen, ctx=True, gates=True, modality="m", tracks=True, ints=True):
    """Random 0/1 dropout masks in the layout of oracle/model.py's `masks` dict.  rows = encoder rows of
    the ints branch (B, or B*T for the track models); context masks are [rows, S, J]."""
    slots = [s for s in ("txt", "vis", "tracks1", "tracks2")
             if (s == "txt" and modality in ("m", "t")) or (s == "vis" and modality in ("m", "v"))
             or (s.startswith("tracks") and tracks)]
    width = sum(J if s in ("txt", "vis") else J // 2 for s in slots)

    def draw(*shape):
        return torch.rand(*shape, generator=gen) >= p
    masks = {}
    if ints or kind == "modalities":
        masks = {("l1", "ints", s): draw(rows, J) for s in slots}
        masks[("cat", "ints")] = draw(rows, width)
    if ctx and kind != "modalities":
        for s in slots:
            masks[("l1", "ctx", s)] = draw(rows, S, J)
        masks[("cat", "ctx")] = draw(rows, width)
        if gates:
            masks[("gate",)] = draw(rows, gate_dim)
    return masks


def run_loss(loss, output, batch):
    """Call a reference loss under torch-1.1 mask semantics."""
    with torch11_masks():
        return loss(output, batch)


def load_dataloader():
    """The reference's mixed_utils.classification_dataloader module (unmodified), imported next to
    mlp.model under the same shims.  Returns (opt, module)."""
    if _state.get("dataloader") is not None:
        return _state["opt"], _state["dataloader"]
    opt, _ = load()
    prefixes = ("utils", "mlp", "mixed_utils", "text_utils", "visual_utils", "resume", "moviegraphs")
    saved_argv, saved_path = sys.argv, list(sys.path)
    shadow = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in prefixes}
    sys.modules.update(_state["modules"])
    sys.argv = ["oracle"]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import mixed_utils.classification_dataloader as ref_dl  # noqa
        import mixed_utils.mixed_features as ref_mf  # noqa
        import utils.util_functions as ref_uf  # noqa
        import utils.evaluation as ref_ev  # noqa
        _state["dataloader"], _state["mixed_features"], _state["util_functions"] = ref_dl, ref_mf, ref_uf
        _state["evaluation"] = ref_ev
        _state["modules"] = {k: v for k, v in sys.modules.items() if k.split(".")[0] in prefixes}
    finally:
        sys.argv, sys.path[:] = saved_argv, saved_path
        for k in list(sys.modules):
            if k.split(".")[0] in prefixes:
                del sys.modules[k]
        sys.modules.update(shadow)
    return opt, ref_dl


def reference_dataset(world_split, world, mode, preset, **overrides):
    """Run the reference's UNMODIFIED MixedFeaturesDataset.__init__ / cache() / init_relships() on a
    synthetic annotation world (lirec_b200/mixed_utils/synthetic_world.py): only the file loaders the
    constructor calls are replaced by functions that hand over the world, and the per-scene feature
    holders are the reference's own MixedFeatures objects with their caches pre-filled (so no .npy
    file is read or written)."""
    import contextlib as _cl
    import io
    opt, dl = load_dataloader()
    mf, uf = _state["mixed_features"], _state["util_functions"]
    set_preset(preset, **overrides)
    opt.text_dim, opt.visual_dim, opt.track_dim = world.text_dim, world.visual_dim, world.track_dim
    opt.mlp_dim = world.text_dim + world.visual_dim + 2 * world.track_dim
    opt.inter_class, opt.merged, opt.feature_type = "all", True, "m"
    opt.multilab_weights = True
    opt.rels = False
    holders = {}

    def make_holder(video_idx, scene_idx, fname):
        h = mf.MixedFeatures.__new__(mf.MixedFeatures)
        h.video_idx, h.scene_idx, h.fname = video_idx, scene_idx, fname
        h.visual = h.textual = None
        h.f_text = h.f_visual = None
        h.cached, h.cached_tracks = {}, {}
        for it in world_split["interactions"]:
            if it.video_descr["movie"] == video_idx and it.video_descr["scene"][0] == scene_idx:
                h.cached[it.id] = world_split["clip_vec"][it.id]
                for p in it.id2names.values():
                    h.cached_tracks[(it.id, p)] = world_split["track_vec"][(it.id, p)]
        holders[(video_idx, scene_idx)] = h
        return h

    patches = dict(
        load_interaction_names=lambda: (world.interaction_names, world.inter2idx),
        load_merged_interactions=lambda: (world.inter2mgd, world.mgd2idx),
        load_set=lambda mode=None: world_split["movie_idxs"],
        load_annotated_inter=lambda movie_idxs=None, inter_class=None: (
            (world_split["interactions"], world_split["rels"], world_split["rels_list"], world_split["rels_opp"])
            if (opt.rels or opt.rels_multitask) else world_split["interactions"]),
        load_iou2_clips=lambda: world.iou2_clips,
        MixedFeatures=make_holder,
        tqdm=lambda x, *a, **k: x,
    )
    saved = {k: getattr(dl, k) for k in patches}
    for k, v in patches.items():
        setattr(dl, k, v)
    try:
        with _cl.redirect_stdout(io.StringIO()):
            ds = dl.MixedFeaturesDataset(mode=mode)
            ds.cache()
            if opt.rels or opt.rels_multitask:
                ds.init_relships()
    finally:
        for k, v in saved.items():
            setattr(dl, k, v)
    return ds
