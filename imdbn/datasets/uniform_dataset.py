"""``from imdbn.datasets.uniform_dataset import create_dataloaders_uniform`` (scripts/train_multimodal.py:11)."""
from multimodal_idbn_b200.datasets import create_dataloaders_uniform  # noqa: F401
