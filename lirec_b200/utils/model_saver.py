"""Keep the best-n checkpoints per metric, with the reference's interface and on-disk layout
(utils/model_saver.py:17-64): `ModelSaver(n=4, path='')`, `check(val)`, `update(val, model, epoch)`,
`save()` writing path/<metric>/v<value>_ep<epoch>.pth.tar, deleting the files of checkpoints that fell out
of the top-n and never rewriting one that is already on disk.  The payload is whatever the loop hands over
({'epoch', 'state_dict', 'optimizer'}, mlp/train.py:84-90) — plain torch.save works on the flat-buffer
parameters."""
import os

import torch

from lirec_b200.utils.util_functions import dir_check


class ModelSaver(object):
    def __init__(self, n=4, path=""):
        self.n = n
        self.path = path
        self.kept = {}        # metric -> {epoch: (value, payload)}, insertion-ordered
        self.on_disk = {}     # metric -> {epoch: file path}
        dir_check(path)

    def _worst(self, key):
        """Epoch of the lowest kept value; among equal values the most recently inserted one (the reference
        scans with `<=`, :48-51)."""
        worst_epoch, worst_val = None, None
        for epoch, (value, _) in self.kept[key].items():
            if worst_val is None or value <= worst_val:
                worst_epoch, worst_val = epoch, value
        return worst_epoch

    def check(self, val):
        """True if any metric of `val` would enter its top-n (reference :31-35)."""
        for key, value in val.items():
            kept = self.kept.get(key, {})
            if len(kept) < self.n:
                return True
            if value > kept[self._worst(key)][0]:
                return True
        return False

    def update(self, val, model, epoch):
        for key, value in val.items():
            kept = self.kept.setdefault(key, {})
            evict = self._worst(key) if len(kept) >= self.n and epoch not in kept else None
            kept[epoch] = (value, model)
            if evict is not None:
                kept.pop(evict)
                self.on_disk.setdefault(key, {}).pop(evict, None)
            assert len(kept) <= self.n

    def save(self):
        for key, kept in self.kept.items():
            folder = os.path.join(self.path, key)
            dir_check(folder)
            files = self.on_disk.setdefault(key, {})
            live = set(files.values())
            for name in os.listdir(folder):                      # evicted checkpoints leave the store
                full = os.path.join(folder, name)
                if full not in live:
                    os.remove(full)
            for epoch, (value, payload) in kept.items():
                if epoch in files:
                    continue                                     # already written
                files[epoch] = os.path.join(folder, "v%.4f_ep%d.pth.tar" % (value, epoch))
                torch.save(payload, files[epoch])
