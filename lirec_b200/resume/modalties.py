"""Entry point: multimodal interaction model (reference: resume/modalties.py:79-100 flag preset)."""
from lirec_b200.resume._common import pipeline, released_checkpoint
from lirec_b200.utils.arg_pars import opt


def resume_modalities():
    opt.resume = True
    opt.test = True
    opt.mod_check = True
    opt.ints = 1
    opt.modality = "m"
    opt.feature_type = "m"
    opt.tracks = True
    opt.tr_maximize = False
    opt.inter_class = "m" if opt.sanity_check else "all"
    opt.log_prefix = ""
    released_checkpoint(opt.data_root + "/models_release/mod_all.pth.tar")
    return pipeline("")


if __name__ == "__main__":
    opt.sanity_check = False
    resume_modalities()
