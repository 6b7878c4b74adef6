"""``imdbn.models.imdbn_bimodal`` of the reference, resolved to the CUDA-backed class."""
from multimodal_idbn_b200 import RBM, iDBN, iMDBN_BiModal  # noqa: F401

__all__ = ["iMDBN_BiModal"]
