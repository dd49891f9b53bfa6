"""Keep the best-n checkpoints per metric (reference: utils/model_saver.py:17-64) — plain torch.save
of {'epoch', 'state_dict', 'optimizer'}, which works unchanged on the flat-buffer parameters."""
import os

import torch

from lirec_b200.utils.util_functions import dir_check


class ModelSaver(object):
    def __init__(self, path, top_n=4):
        self.path, self.top_n = path, top_n
        self.best = {}                       # metric -> sorted list of (value, epoch, save_dict)

    def check(self, values):
        """True if any metric of `values` would enter its top-n."""
        for k, v in values.items():
            lst = self.best.get(k, [])
            if len(lst) < self.top_n or v > lst[-1][0]:
                return True
        return False

    def update(self, values, save_dict, epoch):
        for k, v in values.items():
            lst = self.best.setdefault(k, [])
            lst.append((v, epoch, save_dict))
            lst.sort(key=lambda t: -t[0])
            del lst[self.top_n:]

    def save(self):
        dir_check(self.path)
        for k, lst in self.best.items():
            for v, epoch, sd in lst:
                torch.save(sd, os.path.join(self.path, "%s_%d_%.4f.pth.tar" % (k, epoch, v)))
