"""Derive dims / paths into `opt`, seed the RNGs, force the device (reference:
mixed_utils/update_arg_pars.py:19-73).  The B200 path has no CPU fallback, so the device is 'cuda'
or the call fails; directory creation under opt.data_root is skipped when the root does not exist
(offline / synthetic runs)."""
import os
import random

import numpy as np
import torch
import torch.backends.cudnn as cudnn

from lirec_b200.utils.arg_pars import opt

_PATH_FLAGS = ["dialogs_path", "frame2time_path", "labeled_interactions", "merged_interactions", "annotations",
               "split_path", "intersected", "relships2_15", "relships_opp", "merged_videos", "ftack_ids",
               "ftracks", "orig_res"]


def update(model_name):
    if not torch.cuda.is_available():
        raise RuntimeError("lirec_b200 needs a B200 GPU: torch.cuda.is_available() is False and there is no "
                           "CPU fallback (the reference's CPU path lives in oracle/ for tests only)")
    opt.device = "cuda"
    torch.manual_seed(opt.seed)
    torch.cuda.manual_seed(opt.seed)
    np.random.seed(opt.seed)
    random.seed(opt.seed)
    torch.backends.cudnn.deterministic = True
    cudnn.benchmark = False

    opt.visual_path = opt.data_root + "/features/spat_i3d"
    opt.visual_dim = 2048
    opt.sampling_fr = 0.0625
    opt.bert_model = "bert_base_uncased"
    opt.text_path = opt.data_root + "/features/bert/bert_base"
    opt.text_dim = 768
    opt.text_layers = 12
    if opt.feature_type == "v":
        opt.text_dim = 0
    if opt.feature_type == "t":
        opt.visual_dim = 0
    opt.mlp_dim = opt.visual_dim + opt.text_dim
    if opt.tracks:
        opt.track_dim = opt.visual_dim
        opt.mlp_dim = opt.mlp_dim + opt.track_dim * 2
    opt.model_name = model_name
    if os.path.isdir(opt.data_root):
        os.makedirs(opt.visual_path, exist_ok=True)
    if not getattr(opt, "_paths_expanded", False):
        for k in _PATH_FLAGS:
            setattr(opt, k, opt.data_root + getattr(opt, k))
        opt._paths_expanded = True
    for arg in sorted(vars(opt)):
        print("%s: %s" % (arg, getattr(opt, arg)))
