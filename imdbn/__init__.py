"""Drop-in alias of the reference's ``imdbn`` package: same import paths, classes backed by the
sm_100a kernels of ``multimodal_idbn_b200``."""
