"""Entry point: multi-task interaction + relationship model (reference: resume/int_rels.py:88-115)."""
from lirec_b200.resume._common import pipeline, released_checkpoint
from lirec_b200.utils.arg_pars import opt


def resume_ints_rels():
    opt.resume = True
    opt.test = True
    opt.feature_type = "m"
    opt.tracks = True
    opt.tr_maximize = False
    opt.mod_check = False
    opt.rels_multitask = True
    opt.rels_multi_clip = True
    opt.rels_n_clips = 18
    opt.ints = 1
    opt.gates = 1
    opt.ctx = 1
    opt.lymbda = 1
    opt.inter_class = "m" if opt.sanity_check else "all"
    opt.log_prefix = ""
    released_checkpoint(opt.data_root + "/models_release/int_rel.pth.tar")
    return pipeline("")


if __name__ == "__main__":
    opt.sanity_check = False
    resume_ints_rels()
