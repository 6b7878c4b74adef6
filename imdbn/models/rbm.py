from multimodal_idbn_b200.rbm import RBM  # noqa: F401
