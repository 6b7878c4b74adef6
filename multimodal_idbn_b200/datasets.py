"""GPU-resident dataset and loaders.

The reference imports ``imdbn.datasets.uniform_dataset.create_dataloaders_uniform`` (scripts/train_multimodal.py:11,
96-102; the module itself is absent from the reference tree) and then pays, per minibatch, a host->device copy plus
``.view(B, -1).float()`` (idbn.py:199-200, imdbn.py:554-555).  Here the whole ``.npz`` lives in HBM as flattened fp32
images and one-hot labels; an epoch is ONE on-device permutation of the rows, and every minibatch is a contiguous,
16-byte-aligned VIEW of the permuted matrix -- exactly what the TMA descriptors of the kernels want, no copy at all.

Same surface as the torch objects the model classes touch: ``loader.dataset`` (a ``Subset``-like object with
``.indices`` and ``.dataset``), ``len(loader)``, ``loader.batch_size``, iteration yielding ``(images, labels)``, and
on the base dataset the per-sample lists ``labels / cumArea_list / CH_list / density_list / N_list`` that
``iDBN.__init__`` reads from the validation split (idbn.py:131-137).
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

_IMAGE_KEYS = ("D", "images", "X", "data", "x")
_LABEL_KEYS = ("N_list", "labels", "y", "targets")
_FEATURE_KEYS = ("cumArea_list", "CH_list", "density_list", "N_list")


class DeviceDataset:
    """Images ``[N, D]`` fp32 and one-hot labels ``[N, K]`` fp32 resident on ``device``."""

    def __init__(self, images, labels, device, num_classes: Optional[int] = None, features: Optional[dict] = None,
                 image_shape: Optional[Tuple[int, ...]] = None):
        dev = torch.device(device)
        x = torch.as_tensor(images)
        self.image_shape = tuple(image_shape) if image_shape is not None else tuple(x.shape[1:])
        self.images = x.reshape(x.shape[0], -1).to(device=dev, dtype=torch.float32).contiguous()
        y = torch.as_tensor(labels)
        if y.dim() == 1 or (y.dim() == 2 and y.shape[1] == 1):
            idx = y.reshape(-1).long()
            lo = int(idx.min()) if idx.numel() else 0
            idx = idx - min(lo, 1) if lo >= 1 else idx          # numerosities 1..K -> classes 0..K-1
            K = int(num_classes) if num_classes is not None else (int(idx.max()) + 1 if idx.numel() else 1)
            self.labels = [int(v) for v in torch.as_tensor(labels).reshape(-1).tolist()]
            y = torch.nn.functional.one_hot(idx, K)
        else:
            self.labels = [int(v) for v in y.argmax(dim=1).tolist()]
        self.onehot = y.to(device=dev, dtype=torch.float32).contiguous()
        if self.onehot.shape[0] != self.images.shape[0]:
            raise ValueError("images and labels differ in length")
        n = self.images.shape[0]
        feats = features or {}
        self.cumArea_list = list(feats.get("cumArea_list", [0.0] * n))
        self.CH_list = list(feats.get("CH_list", [0.0] * n))
        self.density_list = list(feats["density_list"]) if "density_list" in feats else None
        self.N_list = list(feats.get("N_list", self.labels))
        self.device = dev

    def __len__(self):
        return self.images.shape[0]

    def __getitem__(self, i):
        return self.images[i].view(self.image_shape), self.onehot[i]

    @classmethod
    def from_npz(cls, path, device, num_classes: Optional[int] = None):
        with np.load(path, allow_pickle=False) as z:
            img_key = next((k for k in _IMAGE_KEYS if k in z.files), None)
            lab_key = next((k for k in _LABEL_KEYS if k in z.files), None)
            if img_key is None or lab_key is None:
                raise KeyError(f"{path}: need one of {_IMAGE_KEYS} and one of {_LABEL_KEYS}, found {z.files}")
            images, labels = z[img_key], z[lab_key]
            feats = {k: z[k].tolist() for k in _FEATURE_KEYS if k in z.files and len(z[k]) == len(images)}
        return cls(torch.from_numpy(np.ascontiguousarray(images)), torch.from_numpy(np.ascontiguousarray(labels)),
                   device, num_classes=num_classes, features=feats)


class DeviceSubset:
    """``torch.utils.data.Subset`` look-alike (``.dataset`` / ``.indices``) over a :class:`DeviceDataset`."""

    def __init__(self, dataset: DeviceDataset, indices: Sequence[int]):
        self.dataset = dataset
        self.indices = [int(i) for i in indices]
        self._idx = torch.as_tensor(self.indices, dtype=torch.long, device=dataset.device)

    def __len__(self):
        return len(self.indices)

    def __getitem__(self, i):
        return self.dataset[self.indices[i]]


class DeviceLoader:
    """Minibatches of a device-resident (sub)set: ``(images [B, *image_shape], labels [B, K])`` device tensors.

    ``shuffle=True``: one ``randperm`` + one row gather per epoch (device generator seeded with ``seed + epoch``), the
    batches are then contiguous views.  No ``drop_last``: the last batch may be ragged, as with the reference's
    loaders.  ``device_resident = True`` tells ``prefetch_to_device`` that there is nothing to stage."""
    device_resident = True

    def __init__(self, dataset, batch_size: int, shuffle: bool = False, seed: int = 0, flat: bool = False):
        self.dataset = dataset
        self.batch_size = int(batch_size)
        self.shuffle = bool(shuffle)
        self.seed = int(seed)
        self.flat = bool(flat)
        self._epoch = 0
        base = dataset.dataset if isinstance(dataset, DeviceSubset) else dataset
        self._base = base
        if isinstance(dataset, DeviceSubset):
            self._x = base.images.index_select(0, dataset._idx)
            self._y = base.onehot.index_select(0, dataset._idx)
        else:
            self._x, self._y = base.images, base.onehot

    def __len__(self):
        return math.ceil(self._x.shape[0] / self.batch_size) if self._x.shape[0] else 0

    def __iter__(self):
        x, y = self._x, self._y
        n = x.shape[0]
        if self.shuffle and n > 1:
            g = torch.Generator(device=x.device)
            g.manual_seed(self.seed + self._epoch)
            perm = torch.randperm(n, device=x.device, generator=g)
            x, y = x.index_select(0, perm), y.index_select(0, perm)
        self._epoch += 1
        shape = self._base.image_shape
        for b0 in range(0, n, self.batch_size):
            xb = x[b0:b0 + self.batch_size]
            yield (xb if self.flat else xb.view(xb.shape[0], *shape)), y[b0:b0 + self.batch_size]


def create_dataloaders_uniform(data_path, data_name, batch_size: int = 128, num_workers: int = 0,
                               multimodal_flag: bool = True, second_modality=None, mnist100_path=None,
                               device=None, splits: Tuple[float, float, float] = (0.8, 0.1, 0.1), seed: int = 0,
                               num_classes: Optional[int] = None):
    """``(train_loader, val_loader, test_loader)`` over ``<data_path>/<data_name>`` (an ``.npz``) with the call
    signature the reference's scripts use (scripts/train_multimodal.py:96-102).  ``num_workers`` is accepted and
    ignored: nothing is loaded per batch.  The split is a fixed seeded permutation; the validation loader's
    ``dataset`` is a subset whose base carries the per-sample feature lists."""
    import os
    dev = torch.device(device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))
    path = os.path.join(str(data_path), str(data_name)) if data_path is not None else str(data_name)
    if not os.path.exists(path) and os.path.exists(path + ".npz"):
        path += ".npz"
    base = DeviceDataset.from_npz(path, dev, num_classes=num_classes)
    n = len(base)
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(int(seed))).tolist()
    n_tr = int(round(splits[0] * n))
    n_va = int(round(splits[1] * n))
    parts = perm[:n_tr], perm[n_tr:n_tr + n_va], perm[n_tr + n_va:]
    train = DeviceLoader(DeviceSubset(base, parts[0]), batch_size, shuffle=True, seed=seed)
    val = DeviceLoader(DeviceSubset(base, parts[1]), batch_size, shuffle=False)
    test = DeviceLoader(DeviceSubset(base, parts[2]), batch_size, shuffle=False)
    return train, val, test
