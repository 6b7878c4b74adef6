from multimodal_idbn_b200.imdbn_logging import label_clamped_trajectory  # noqa: F401
