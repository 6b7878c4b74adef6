"""``imdbn.models`` of the reference (imdbn/models/__init__.py:1-35), resolved to the CUDA-backed
classes, including the legacy ``src.classes.*`` aliases old Groundeep pickles need."""
import sys
from types import ModuleType

from multimodal_idbn_b200 import RBM, iDBN, iMDBN

__all__ = ["RBM", "iDBN", "iMDBN"]

_this = sys.modules[__name__]
_src = sys.modules.setdefault("src", ModuleType("src"))
_classes = sys.modules.setdefault("src.classes", ModuleType("src.classes"))
_src.classes = _classes
for _legacy in ("rbm_model", "dbn_model", "gdbn_model"):
    setattr(_classes, _legacy, _this)
    sys.modules.setdefault(f"src.classes.{_legacy}", _this)
