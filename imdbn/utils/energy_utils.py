"""``imdbn.utils.energy_utils`` of the reference, resolved to the CUDA-backed functions."""
from multimodal_idbn_b200.energy_utils import (  # noqa: F401
    rbm_free_energy, class_free_energies, _deterministic_img2txt_step, trace_single_img2txt,
    trace_batch_img2txt, pick_fixed_val_case, pick_val_case)
