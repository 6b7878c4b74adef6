"""multimodal_idbn_b200 -- B200-native (sm_100a) implementation of the RBM contrastive-divergence /
conditional-Gibbs hot path of francesco-cal98/multimodal-idbn, behind the reference's Python API.

    from multimodal_idbn_b200 import RBM, iDBN, iMDBN, iMDBN_BiModal     # or, as a drop-in:
    from imdbn.models import RBM, iDBN, iMDBN

The arithmetic lives in ``libimdbn_b200.so`` (hand-written CUDA, C ABI in ``include/imdbn_b200.h``);
there is no CPU fallback.
"""
from ._lib import LIB_PATH, get_precision, load_library, set_precision, total_launches
from .rbm import RBM, rbm_free_energy, random_field
from .idbn import iDBN, prefetch_to_device
from .imdbn import iMDBN
from .imdbn_bimodal import iMDBN_BiModal
from . import dist
from . import datasets

# Checkpoints must cross-load with the reference: classes pickle under the reference's
# module paths, which the ``imdbn`` alias package at the repository root resolves to these classes.
RBM.__module__ = "imdbn.models.rbm"
iDBN.__module__ = "imdbn.models.idbn"
iMDBN.__module__ = "imdbn.models.imdbn"
iMDBN_BiModal.__module__ = "imdbn.models.imdbn_bimodal"

__all__ = ["RBM", "iDBN", "iMDBN", "iMDBN_BiModal", "rbm_free_energy", "random_field", "prefetch_to_device", "dist", "datasets",
           "set_precision", "get_precision", "load_library", "total_launches", "LIB_PATH"]
