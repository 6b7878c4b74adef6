from multimodal_idbn_b200.conditional_steps import (  # noqa: F401
    _gibbs_conditional_step, _steps_stats, build_or_get_fixed_val_panel, log_cross_case, pick_fixed_val_case,
    run_and_log_cross_fixed_case, run_and_log_cross_panel, run_and_log_z_mismatch_check, run_cross_panel,
    trace_img2txt_cross, trace_txt2img_cross, z_mismatch_stats)
