"""Entry point: interaction + character-pair detection with the weakly supervised track-assignment
loss (reference: resume/int_ch.py:77-117)."""
from lirec_b200.resume._common import pipeline, released_checkpoint
from lirec_b200.utils.arg_pars import opt


def resume_max_tracks():
    opt.resume = True
    opt.test = True
    opt.visdom = False
    opt.tr_maximize = True
    opt.feature_type = "m"
    opt.tracks = True
    opt.mod_check = False
    opt.ints = 1
    opt.ctx = 0
    opt.gates = 0
    opt.rels_multitask = False
    opt.rels_multi_clip = False
    opt.inter_class = "m" if opt.sanity_check else "all"
    opt.log_prefix = ""
    name = "gt_int_ch_sum_max" if opt.tr_correct else "weak_int_ch_sum_max"
    released_checkpoint(opt.data_root + "/models_release/%s.pth.tar" % name)
    return pipeline("")


if __name__ == "__main__":
    opt.tr_correct = False          # True = ground-truth supervised assignment, False = weak training
    opt.sanity_check = False
    resume_max_tracks()
