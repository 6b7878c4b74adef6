"""Linear probes on the learned representations, device resident.

Mirror of the reference's ``imdbn/utils/probe_utils.py`` for the compute side: the validation embeddings come from
``represent`` (the CUDA up passes) and STAY in HBM, the quantile binning runs on the device, the full-batch softmax
probe (``nn.Linear`` + AdamW + early stopping on the validation loss, probe_utils.py:195-263) trains on the device
without the NumPy round trip of the reference, and the confusion matrix is one ``bincount``.  Same function names,
arguments and return values; what the reference renders (W&B tables, bar charts, CSV files, probe_utils.py:268-330,
420-433) is out of scope: scalars are logged when the model carries a ``wandb_run``, and every orchestrator RETURNS
its summary so that callers and tests can read it.

The stratified split keeps the reference's host procedure bit for bit (``random.Random(seed).shuffle`` per class, in
``torch.unique`` order, probe_utils.py:170-189): the index sets are part of the result."""
from __future__ import annotations

import random
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

_FEATURE_NAMES = (("cum_area", ("Cumulative Area", "cum_area")),
                  ("convex_hull", ("Convex Hull", "convex_hull", "convexhull")),
                  ("labels", ("Labels", "labels")),
                  ("density", ("Density", "density")))


def _norm(name: str) -> str:
    return name.lower().replace(" ", "").replace("_", "")


def _features_of(model, n: int, device) -> Dict[str, torch.Tensor]:
    """``model.features`` under the reference's canonical keys, 1-D fp32 (one-hot matrices become indices) on
    ``device`` (probe_utils.py:31-80)."""
    src = getattr(model, "features", None)
    if src is None:
        raise RuntimeError("model.features is required")
    by_norm = {_norm(k): k for k in src.keys()}
    feats: Dict[str, torch.Tensor] = {}
    for key, candidates in _FEATURE_NAMES:
        found = next((by_norm[_norm(c)] for c in candidates if _norm(c) in by_norm), None)
        if found is None:
            continue
        t = torch.as_tensor(src[found])
        if t.ndim == 2:
            t = torch.argmax(t, dim=1)
        t = t.reshape(-1).to(device=device, dtype=torch.float32)
        if t.numel() != n:
            raise RuntimeError(f"Feature '{key}' length mismatch: {t.numel()} vs embeddings {n}.")
        feats[key] = t
    return feats


@torch.no_grad()
def compute_val_embeddings_and_features(model, upto_layer: Optional[int] = None) -> Tuple[torch.Tensor, dict]:
    """Embeddings of the whole validation loader, [N, D] on the model's device (probe_utils.py:20-80)."""
    assert model.val_loader is not None, "val_loader is None."
    parts = []
    for batch_data, batch_labels in model.val_loader:
        x = batch_labels if getattr(model, "text_flag", False) else batch_data
        x = x.to(model.device).reshape(x.shape[0], -1).float()
        parts.append(model.represent(x) if upto_layer is None else model.represent(x, upto_layer=upto_layer))
    E = torch.cat(parts, dim=0)
    return E, _features_of(model, E.shape[0], E.device)


@torch.no_grad()
def compute_joint_embeddings_and_features(model) -> Tuple[torch.Tensor, dict]:
    """Joint (image, label) embeddings of the validation loader (probe_utils.py:84-145)."""
    assert model.val_loader is not None, "val_loader is None."
    parts = [model.represent((img.to(model.device), lab.to(model.device))) for img, lab in model.val_loader]
    if not parts:
        return torch.empty(0), {}
    E = torch.cat(parts, dim=0)
    return E, _features_of(model, E.shape[0], E.device)


def make_bin_labels(values: torch.Tensor, n_bins: int = 5):
    """Quantile binning: labels in 0..n_bins-1 and the n_bins+1 edges; equal edges are separated by 1e-6
    (probe_utils.py:151-166).  Runs where ``values`` lives."""
    qs = torch.linspace(0, 1, steps=n_bins + 1, device=values.device)
    edges = torch.quantile(values, qs, interpolation="linear")
    e = edges.tolist()
    for k in range(1, len(e)):
        if e[k] <= e[k - 1]:
            e[k] = e[k - 1] + 1e-6
    edges = torch.tensor(e, dtype=edges.dtype, device=values.device)
    labels = torch.bucketize(values, edges[1:-1].contiguous(), right=False)
    return labels, edges


def _format_bin_names(edges: torch.Tensor, precision: int = 4) -> List[str]:
    def fmt(v: float) -> str:
        return f"{v:.{precision}f}".rstrip("0").rstrip(".")
    e = [float(v) for v in edges.tolist()]
    return [f"{fmt(lo)}-{fmt(hi)}" for lo, hi in zip(e[:-1], e[1:])]


def stratified_split(labels: torch.Tensor, test_size: float = 0.2, rng_seed: int = 42):
    """Per-class shuffled split over ALL samples, at least one test and one training sample per class that has two
    (probe_utils.py:170-189).  Host procedure kept exactly: the index lists are reproducible across implementations."""
    rng = random.Random(rng_seed)
    host = labels.detach().cpu()
    train_idx: List[int] = []
    test_idx: List[int] = []
    for c in torch.unique(host).tolist():
        members = (host == c).nonzero(as_tuple=True)[0].tolist()
        rng.shuffle(members)
        if len(members) <= 1:
            test_idx.extend(members)
            continue
        n_test = min(max(1, int(round(len(members) * test_size))), len(members) - 1)
        test_idx.extend(members[:n_test])
        train_idx.extend(members[n_test:])
    return train_idx, test_idx


def train_linear_classifier(X_train, y_train, X_val, y_val, device, n_classes: int, max_steps: int = 1000,
                            lr: float = 1e-2, weight_decay: float = 0.0, patience: int = 20, min_delta: float = 0.0):
    """Full-batch linear probe with early stopping on the validation loss (probe_utils.py:195-263).  Inputs may be
    device tensors (no copy) or arrays.  The validation loss of every step stays on the device; the early-stopping
    decision needs it on the host, so it is read in blocks of ``patience`` steps and the loop is replayed from the
    block's snapshots -- same stopping step and same selected parameters as the step-by-step loop.
    Returns (accuracy at the best step, y_true list, y_pred list)."""
    device = torch.device(device)
    Xtr = torch.as_tensor(X_train, dtype=torch.float32).to(device)
    ytr = torch.as_tensor(y_train).to(device=device, dtype=torch.long)
    Xva = torch.as_tensor(X_val, dtype=torch.float32).to(device)
    yva = torch.as_tensor(y_val).to(device=device, dtype=torch.long)
    model = nn.Linear(Xtr.shape[1], n_classes).to(device)        # (created on the host like the reference's: same init draw)
    opt = torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=weight_decay)

    best_loss, best_state, no_improve = float("inf"), None, 0
    block = max(1, int(patience))
    step, stop = 0, False
    while step < max_steps and not stop:
        n = min(block, max_steps - step)
        losses = torch.empty(n, device=device)
        states = []
        for i in range(n):
            opt.zero_grad(set_to_none=True)
            F.cross_entropy(model(Xtr), ytr).backward()
            opt.step()
            with torch.no_grad():
                losses[i] = F.cross_entropy(model(Xva), yva)
                states.append((model.weight.detach().clone(), model.bias.detach().clone()))
        for i, v_loss in enumerate(losses.tolist()):              # ONE device->host read per block
            if v_loss < best_loss - min_delta:
                best_loss, best_state, no_improve = v_loss, states[i], 0
            else:
                no_improve += 1
                if no_improve >= patience:
                    stop = True
                    break
        step += n
    with torch.no_grad():
        if best_state is not None:
            model.weight.copy_(best_state[0])
            model.bias.copy_(best_state[1])
        preds = torch.argmax(model(Xva), dim=1)
        acc = (preds == yva).float().mean().item()
    return acc, yva.tolist(), preds.tolist()


def confusion_matrix(y_true, y_pred, n_classes: int) -> torch.Tensor:
    """[true, pred] counts (the table of probe_utils.py:268-281, without the DataFrame)."""
    t = torch.as_tensor(y_true, dtype=torch.long)
    p = torch.as_tensor(y_pred, dtype=torch.long)
    ok = (t >= 0) & (t < n_classes) & (p >= 0) & (p < n_classes)
    return torch.bincount(t[ok] * n_classes + p[ok], minlength=n_classes * n_classes).reshape(n_classes, n_classes)


@torch.no_grad()
def pca_project(E: torch.Tensor, n_components: int = 2) -> Tuple[torch.Tensor, torch.Tensor]:
    """Principal-component projection of the embeddings on the device: what ``sklearn.decomposition.PCA(n).
    fit_transform`` returns for the figures of idbn.py:262-283, up to the sign of each component, and the explained
    variance ratios.  Exact SVD of the centred matrix (N x D with D <= a few thousand)."""
    X = E.to(torch.float32)
    X = X - X.mean(dim=0, keepdim=True)
    U, S, _ = torch.linalg.svd(X, full_matrices=False)
    k = min(int(n_components), S.numel())
    var = S * S
    return U[:, :k] * S[:k], var[:k] / var.sum().clamp_min(1e-30)


def _prepare_targets(feats: dict, mkey: str, n_bins: int):
    y, edges = make_bin_labels(feats[mkey].to(torch.float32), n_bins=n_bins)
    return y.long(), n_bins, edges, _format_bin_names(edges, precision=4)


def _probe(model, E: torch.Tensor, feats: dict, epoch: int, prefix: Optional[str], n_bins, test_size, steps, lr,
           rng_seed, patience, min_delta) -> Dict[str, dict]:
    run = getattr(model, "wandb_run", None)
    out: Dict[str, dict] = {}
    targets = ["cum_area", "convex_hull", "labels"] + (["density"] if "density" in feats else [])
    for mkey in targets:
        if mkey not in feats:
            continue
        y, n_classes, edges, bin_names = _prepare_targets(feats, mkey, n_bins)
        name = f"{prefix}/{mkey}" if prefix else mkey
        train_idx, test_idx = stratified_split(y, test_size=test_size, rng_seed=rng_seed)
        if not train_idx or not test_idx:
            if run is not None:
                run.log({f"probe/{name}/warn_empty_split/acc": 0.0, "epoch": epoch})
            continue
        tr = torch.tensor(train_idx, device=E.device)
        te = torch.tensor(test_idx, device=E.device)
        acc, y_true, y_pred = train_linear_classifier(E[tr], y[tr], E[te], y[te], device=model.device,
                                                      n_classes=n_classes, max_steps=steps, lr=lr, weight_decay=0.0,
                                                      patience=patience, min_delta=min_delta)
        out[name] = {"acc": acc, "confusion": confusion_matrix(y_true, y_pred, n_classes), "bin_names": bin_names,
                     "bin_edges": edges.detach().cpu()}
        if run is not None:
            run.log({f"probe/{name}/acc": acc, "epoch": epoch})
    return out


def log_linear_probe(model, epoch: int, n_bins: int = 5, test_size: float = 0.2, steps: int = 1000, lr: float = 1e-2,
                     rng_seed: int = 42, patience: int = 20, min_delta: float = 0.0, save_csv: bool = True,
                     upto_layer: Optional[int] = None, layer_tag: Optional[str] = None) -> Dict[str, dict]:
    """Probes on cum_area / convex_hull / labels (/ density), all binned into ``n_bins`` levels
    (probe_utils.py:344-433).  ``save_csv`` is accepted for signature compatibility (no files are written)."""
    E, feats = compute_val_embeddings_and_features(model, upto_layer=upto_layer)
    return _probe(model, E, feats, epoch, layer_tag, n_bins, test_size, steps, lr, rng_seed, patience, min_delta)


def log_joint_linear_probe(model, epoch: int, n_bins: int = 5, test_size: float = 0.2, steps: int = 1000,
                           lr: float = 1e-2, rng_seed: int = 42, patience: int = 20, min_delta: float = 0.0,
                           save_csv: bool = False, metric_prefix: str = "joint") -> Dict[str, dict]:
    """The same on the joint embeddings of an iMDBN (probe_utils.py:436-510)."""
    E, feats = compute_joint_embeddings_and_features(model)
    if E.numel() == 0:
        return {}
    return _probe(model, E, feats, epoch, metric_prefix, n_bins, test_size, steps, lr, rng_seed, patience, min_delta)
