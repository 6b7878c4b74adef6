from multimodal_idbn_b200.conditional_steps import (  # noqa: F401
    _gibbs_conditional_step, run_cross_panel, trace_img2txt_cross, trace_txt2img_cross)
