"""Drop-in ``iMDBN_BiModal`` (reference ``imdbn/models/imdbn_bimodal.py:422-1076``): two image iDBNs
(numerosity, MNIST-100) joined by a stack of RBMs over ``[z_mod1 (+) z_mod2]``, on the CUDA kernels.

Same kernels as ``iMDBN`` with different flags: ``train_epoch_clamped(CD=3,
sample_h=True)`` in both clamping directions (bimodal.py:755-820), a multi-layer joint stack trained per
batch (:785-792), ``conditional_gibbs(sample_h=True)`` in both directions (:673-686).  Per-batch metrics are
accumulated on the device and read back once per epoch (the reference issues two ``.item()`` per batch,
:830-831).  W&B / PCA / probe / trajectory rendering (:43-420, :836-1015) is out of scope: ``metrics_history``
receives the per-epoch numbers the reference logs.
"""
from __future__ import annotations

import datetime
import pickle
from typing import Any, Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .idbn import iDBN, _flat, prefetch_to_device
from .rbm import RBM


class iMDBN_BiModal(nn.Module):
    """See the reference docstring (imdbn_bimodal.py:422-435)."""

    WARMUP_EPOCHS = 8            # imdbn_bimodal.py:734

    def __init__(self, layer_sizes_mod1: list, layer_sizes_mod2: list, joint_layer_sizes, params: Optional[dict] = None,
                 dataloader=None, val_loader=None, device=None, wandb_run=None, logging_cfg: Optional[dict] = None):
        super().__init__()
        self.params = params or {}
        self.device = torch.device(device) if device is not None else torch.device(
            "cuda" if torch.cuda.is_available() else "cpu")
        self.dataloader = dataloader
        self.val_loader = val_loader
        self.wandb_run = wandb_run
        self.logging_cfg = logging_cfg or {}
        self.mod1_dbn = iDBN(layer_sizes=layer_sizes_mod1, params=self.params, dataloader=None, val_loader=None,
                             device=self.device, wandb_run=self.wandb_run)
        self.mod2_dbn = iDBN(layer_sizes=layer_sizes_mod2, params=self.params, dataloader=None, val_loader=None,
                             device=self.device, wandb_run=self.wandb_run)
        self.Dz_mod1 = int(self.mod1_dbn.layers[-1].num_hidden)
        self.Dz_mod2 = int(self.mod2_dbn.layers[-1].num_hidden)
        self._build_joint(joint_layer_sizes)
        self.joint_cd = int(self.params.get("JOINT_CD", self.params.get("CD", 1)))          # :505
        self.cross_steps = int(self.params.get("CROSS_GIBBS_STEPS", 50))

        try:                                                                              # :509-515
            vb_mod1, vb_mod2 = next(iter(val_loader))
            self.validation_mod1 = vb_mod1[:8].to(self.device)
            self.validation_mod2 = vb_mod2[:8].to(self.device)
        except Exception:
            self.validation_mod1 = None
            self.validation_mod2 = None

        self.features = None                                                              # :518-540
        try:
            if hasattr(val_loader.dataset, "indices"):
                indices, base = val_loader.dataset.indices, val_loader.dataset.dataset
            else:
                base = val_loader.dataset
                indices = range(len(base))
            self.features = {
                "Cumulative Area": torch.tensor([base.cumArea_list[i] for i in indices], dtype=torch.float32),
                "Convex Hull": torch.tensor([base.CH_list[i] for i in indices], dtype=torch.float32),
                "Labels": torch.tensor([base.labels[i] for i in indices], dtype=torch.float32),
            }
            density_src = getattr(base, "density_list", None)
            if density_src is not None:
                self.features["Density"] = torch.tensor([density_src[i] for i in indices], dtype=torch.float32)
        except Exception:
            self.features = None

        joint_sizes = joint_layer_sizes if isinstance(joint_layer_sizes, list) else [joint_layer_sizes]
        self.arch_str = (f"MOD1{'-'.join(map(str, layer_sizes_mod1))}_MOD2{'-'.join(map(str, layer_sizes_mod2))}"
                         f"_JOINT{'-'.join(map(str, joint_sizes))}")
        self.metrics_history = []

    def _build_joint(self, joint_layer_sizes) -> None:
        """Stack of RBMs over the concatenated latents, no softmax groups (imdbn_bimodal.py:543-575)."""
        if isinstance(joint_layer_sizes, int):
            joint_layer_sizes = [joint_layer_sizes]
        self.joint_layers = []
        cur = self.Dz_mod1 + self.Dz_mod2
        for hidden in joint_layer_sizes:
            rbm = RBM(num_visible=cur, num_hidden=int(hidden),
                      learning_rate=self.params.get("JOINT_LEARNING_RATE", self.params.get("LEARNING_RATE", 0.1)),
                      weight_decay=self.params.get("WEIGHT_PENALTY", 0.0001),
                      momentum=self.params.get("INIT_MOMENTUM", 0.5),
                      dynamic_lr=self.params.get("LEARNING_RATE_DYNAMIC", True),
                      final_momentum=self.params.get("FINAL_MOMENTUM", 0.95),
                      softmax_groups=[]).to(self.device)
            self.joint_layers.append(rbm)
            cur = int(hidden)
        self.joint_rbm = self.joint_layers[0]
        self.num_joint_layers = len(self.joint_layers)

    # ------------------------------------------------------------------ loaders (:577-615)
    def load_pretrained_mod1_dbn(self, path: str) -> bool:
        return self._load_pretrained_dbn(self.mod1_dbn, path, "mod1")

    def load_pretrained_mod2_dbn(self, path: str) -> bool:
        return self._load_pretrained_dbn(self.mod2_dbn, path, "mod2")

    def _load_pretrained_dbn(self, dbn: iDBN, path: str, name: str) -> bool:
        try:
            with open(path, "rb") as f:
                obj = pickle.load(f)
        except Exception as e:                                   # noqa: BLE001 - the reference reports and returns False
            print(f"[load_pretrained_{name}_dbn] error: {e}")
            return False
        if isinstance(obj, dict) and "layers" in obj:
            dbn.layers = obj["layers"]
        elif hasattr(obj, "layers"):
            dbn.layers = obj.layers
        else:
            print(f"[load_pretrained_{name}_dbn] unrecognized format")
            return False
        for rbm in dbn.layers:
            rbm.to(self.device)
            rbm.W_m = torch.zeros_like(rbm.W.data)
            rbm.hb_m = torch.zeros_like(rbm.hid_bias.data)
            rbm.vb_m = torch.zeros_like(rbm.vis_bias.data)
            if not hasattr(rbm, "softmax_groups"):
                rbm.softmax_groups = []
        print(f"[load_pretrained_{name}_dbn] loaded from {path}")
        return True

    # ------------------------------------------------------------------ hot paths
    @torch.no_grad()
    def init_joint_bias_from_data(self, n_batches: int = 10) -> None:
        """vis_bias of the first joint layer = logit of the mean latent of each modality (:617-646)."""
        sum1 = sum2 = None
        n = 0
        for b, (mod1, mod2) in enumerate(self.dataloader):
            if b >= n_batches:
                break
            z1 = self.mod1_dbn.represent(_flat(mod1, self.device))
            z2 = self.mod2_dbn.represent(_flat(mod2, self.device))
            sum1 = z1.sum(0) if sum1 is None else sum1 + z1.sum(0)
            sum2 = z2.sum(0) if sum2 is None else sum2 + z2.sum(0)
            n += z1.size(0)
        if n == 0:
            return
        m1 = (sum1 / n).clamp(1e-4, 1 - 1e-4)
        m2 = (sum2 / n).clamp(1e-4, 1 - 1e-4)
        vb = self.joint_layers[0].vis_bias
        vb.data[:self.Dz_mod1] = torch.log(m1) - torch.log1p(-m1)
        vb.data[self.Dz_mod1:] = torch.log(m2) - torch.log1p(-m2)

    def _clamp(self, z: torch.Tensor, first: bool) -> Tuple[torch.Tensor, torch.Tensor]:
        """(v_known, known_mask) with one modality's latents clamped (:655-660, :675-678)."""
        B, Dz1 = z.size(0), self.Dz_mod1
        vk = torch.zeros(B, Dz1 + self.Dz_mod2, device=self.device)
        km = torch.zeros_like(vk)
        if first:
            vk[:, :Dz1] = z
            km[:, :Dz1] = 1.0
        else:
            vk[:, Dz1:] = z
            km[:, Dz1:] = 1.0
        return vk, km

    @torch.no_grad()
    def _cross_reconstruct(self, z_mod1: torch.Tensor, z_mod2: torch.Tensor,
                           steps: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """(mod1_from_mod2, mod2_from_mod1): stochastic-h conditional Gibbs in both directions through the
        first joint layer, decoded by the other modality's iDBN (:648-694)."""
        if steps is None:
            steps = self.cross_steps
        Dz1 = self.Dz_mod1
        vk, km = self._clamp(z_mod1, True)
        z2_from_1 = self.joint_rbm.conditional_gibbs(vk, km, n_steps=steps, sample_h=True, sample_v=False)[:, Dz1:]
        vk, km = self._clamp(z_mod2, False)
        z1_from_2 = self.joint_rbm.conditional_gibbs(vk, km, n_steps=steps, sample_h=True, sample_v=False)[:, :Dz1]
        return self.mod1_dbn.decode(z1_from_2), self.mod2_dbn.decode(z2_from_1)

    @torch.no_grad()
    def represent(self, batch: Tuple[torch.Tensor, torch.Tensor]) -> torch.Tensor:
        """Joint representation through all joint layers (:696-709)."""
        mod1, mod2 = batch
        h = torch.cat([self.mod1_dbn.represent(_flat(mod1, self.device)),
                       self.mod2_dbn.represent(_flat(mod2, self.device))], dim=1)
        for rbm in self.joint_layers:
            h = rbm.forward(h)
        return h

    def train_joint(self, epochs: int, log_every: int = 5, log_every_pca: int = 25, log_every_probe: int = 10,
                    log_every_trajectory: int = 50):
        """Warm-up: alternating clamped CD-3 on the first joint layer; main phase: CD on every joint layer
        plus both clamped directions; cross-modal MSE of every batch (:711-962)."""
        print(f"[iMDBN_BiModal] joint training: {self.num_joint_layers} layers, {epochs} epochs total")
        self.init_joint_bias_from_data(n_batches=10)
        dev = self.device
        aux = int(self.params.get("JOINT_AUX_COND_STEPS", 30))                             # :737
        clamp_kw = dict(CD=3, cond_init_steps=aux, sample_h=True, sample_v=False, aux_lr_mult=0.3,
                        use_noisy_init=True)
        first = self.joint_layers[0]
        for epoch in range(int(epochs)):
            cd_losses = []
            acc = torch.zeros(2, device=dev, dtype=torch.float64)
            n_seen = 0
            for mod1, mod2 in prefetch_to_device(self.dataloader, dev):
                v1, v2 = _flat(mod1, dev), _flat(mod2, dev)
                B = v1.size(0)
                z1 = self.mod1_dbn.represent(v1)
                z2 = self.mod2_dbn.represent(v2)
                vk1, km1 = self._clamp(z1, True)
                vk2, km2 = self._clamp(z2, False)
                if epoch < self.WARMUP_EPOCHS:
                    for _ in range(2):                                                     # :753-778
                        first.train_epoch_clamped(vk1, km1, epoch, epochs, **clamp_kw)
                        first.train_epoch_clamped(vk2, km2, epoch, epochs, **clamp_kw)
                else:
                    cur = torch.cat([z1, z2], dim=1)                                       # :781-792
                    for li, rbm in enumerate(self.joint_layers):
                        loss, cur = rbm.train_epoch_fwd(cur, epoch, epochs, CD=self.joint_cd)
                        if li == 0:
                            cd_losses.append(loss)
                    first.train_epoch_clamped(vk1, km1, epoch, epochs, reclamp_negative=False, **clamp_kw)
                    first.train_epoch_clamped(vk2, km2, epoch, epochs, reclamp_negative=False, **clamp_kw)
                rec1, rec2 = self._cross_reconstruct(z1, z2, steps=self.cross_steps)        # :822-831
                acc += torch.stack([F.mse_loss(rec1.view_as(v1), v1, reduction="sum").double(),
                                    F.mse_loss(rec2.view_as(v2), v2, reduction="sum").double()])
                n_seen += B
            if n_seen > 0:
                s1, s2 = (float(x) for x in acc.cpu())
                rec = {"epoch": epoch,
                       "cross_modality/mod1_mse": s1 / (n_seen * self.mod1_dbn.layers[0].num_visible),
                       "cross_modality/mod2_mse": s2 / (n_seen * self.mod2_dbn.layers[0].num_visible)}
                if cd_losses:
                    rec["joint/cd_loss"] = float(torch.stack(cd_losses).mean())
                self.metrics_history.append(rec)
                if self.wandb_run:
                    self.wandb_run.log(rec)
            if epoch % max(1, int(log_every)) == 0:
                self._log_snapshots(epoch)
        print("[iMDBN_BiModal] joint training finished.")

    def _log_snapshots(self, epoch: int, num: int = 8):
        """W&B snapshot rendering (:964-1015): out of scope for the CUDA path."""
        return None

    # ------------------------------------------------------------------ checkpoints (:1017-1076)
    def save_model(self, path: str):
        for l in list(self.mod1_dbn.layers) + list(self.mod2_dbn.layers) + list(self.joint_layers):
            l.sync_momenta()
        payload = {
            "mod1_dbn": self.mod1_dbn, "mod2_dbn": self.mod2_dbn, "joint_layers": self.joint_layers,
            "num_joint_layers": self.num_joint_layers, "Dz_mod1": self.Dz_mod1, "Dz_mod2": self.Dz_mod2,
            "params": self.params, "arch_str": self.arch_str, "features": self.features,
            "metadata": {"saved_at": datetime.datetime.now().isoformat(), "model_type": "iMDBN_BiModal",
                         "architecture": self.arch_str},
        }
        with open(path, "wb") as f:
            pickle.dump(payload, f)
        print(f"[iMDBN_BiModal] Model saved to {path}")
        print(f"[iMDBN_BiModal] Architecture: {self.arch_str}")

    @staticmethod
    def load_model(path: str, device=None) -> Dict[str, Any]:
        if device is None:
            device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        with open(path, "rb") as f:
            payload = pickle.load(f)
        for key in ("mod1_dbn", "mod2_dbn"):
            if key in payload:
                for rbm in payload[key].layers:
                    rbm.to(device)
        if "joint_layers" in payload:
            for rbm in payload["joint_layers"]:
                rbm.to(device)
        elif "joint_rbm" in payload:                             # old single-RBM format
            payload["joint_rbm"].to(device)
            payload["joint_layers"] = [payload["joint_rbm"]]
            payload["num_joint_layers"] = 1
        print(f"[iMDBN_BiModal] Model loaded from {path}")
        if "arch_str" in payload:
            print(f"[iMDBN_BiModal] Architecture: {payload['arch_str']}")
        return payload
