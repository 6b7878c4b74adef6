"""The reference's single-file copy of the model classes (what ``imdbn.models`` exports there);
pickles written by it name this module."""
from multimodal_idbn_b200 import RBM, iDBN, iMDBN  # noqa: F401
