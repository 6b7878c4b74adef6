from multimodal_idbn_b200.rbm import rbm_free_energy  # noqa: F401
