"""``imdbn.datasets`` of the reference's scripts (absent from the reference tree): the GPU-resident loaders."""
from multimodal_idbn_b200.datasets import (  # noqa: F401
    DeviceDataset, DeviceLoader, DeviceSubset, create_dataloaders_uniform)
