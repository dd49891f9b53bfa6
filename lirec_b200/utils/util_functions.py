"""The few helpers of the reference's utils/util_functions.py that the loops touch: `Averaging`
(:23-38), `dir_check`, `load_model` / `load_optimizer` (:274-291).  Annotation parsing (the other
~550 lines) is dataset plumbing outside the hot path and is not reimplemented."""
import os

import torch

from lirec_b200.utils.arg_pars import opt


class Averaging(object):
    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = 0.0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count if self.count else 0.0


def dir_check(path):
    os.makedirs(path, exist_ok=True)


def load_model(name=""):
    """state_dict of a checkpoint in the reference's format {'epoch','state_dict','optimizer'}."""
    ckpt = torch.load(opt.resume_str, map_location="cpu")
    print("loaded model %s (epoch %s)" % (opt.resume_str, ckpt.get("epoch")))
    return ckpt["state_dict"]


def load_optimizer():
    return torch.load(opt.resume_str, map_location="cpu")["optimizer"]
