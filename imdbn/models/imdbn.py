from multimodal_idbn_b200.imdbn import iMDBN  # noqa: F401
from multimodal_idbn_b200.idbn import iDBN  # noqa: F401
from multimodal_idbn_b200.rbm import RBM  # noqa: F401
