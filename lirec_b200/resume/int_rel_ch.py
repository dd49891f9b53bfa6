"""Entry point: the full model — interactions, relationships and character pairs, data-parallel over
clips when launched under torchrun with --dp 1 (reference: resume/int_rel_ch.py:87-124)."""
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
    opt.ctx = 1
    opt.rels_multitask = True
    opt.rels_multi_clip = True
    opt.gates = 1
    opt.rels_n_clips = 18
    opt.inter_class = "m" if opt.sanity_check else "all"
    opt.log_prefix = ""
    name = "gt_int_rel_ch_sum_max" if opt.tr_correct else "weak_int_rel_ch_sum_max"
    released_checkpoint(opt.data_root + "/models_release/%s.pth.tar" % name)
    return pipeline("")


if __name__ == "__main__":
    opt.tr_correct = False          # True = ground-truth supervised assignment, False = weak training
    opt.sanity_check = False
    resume_max_tracks()
