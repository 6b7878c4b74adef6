"""``imdbn.utils.probe_utils`` of the reference, resolved to the device-resident implementation."""
from multimodal_idbn_b200.probe_utils import (  # noqa: F401
    compute_val_embeddings_and_features, compute_joint_embeddings_and_features, make_bin_labels, _format_bin_names,
    stratified_split, train_linear_classifier, _prepare_targets, log_linear_probe, log_joint_linear_probe,
    confusion_matrix, pca_project)
