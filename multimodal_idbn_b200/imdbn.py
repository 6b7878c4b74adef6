"""Drop-in ``iMDBN`` (reference ``imdbn/models/imdbn.py``): image iDBN + joint RBM over
``[z_img (+) one-hot label]`` with a 32-way softmax group, on the CUDA kernels.

Hot paths re-hosted here: ``init_joint_bias_from_data`` (imdbn.py:216-292), ``train_joint``
(508-639), ``_cross_reconstruct`` (386-488), ``represent`` (490-506), the loaders and the dual-format
``save_model`` pickle (815-934).  Per-batch metrics are accumulated on the device and read back once
per epoch (the reference issues four ``.item()`` per batch, imdbn.py:635-638).  W&B snapshot
rendering (``_log_snapshots``, 714-813) is out of scope and is a no-op.
"""
from __future__ import annotations

import ctypes as C
import datetime
import os
import pickle
from typing import Any, Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L
from .idbn import iDBN, _flat, prefetch_to_device
from .rbm import RBM


class iMDBN(nn.Module):
    """See the reference docstring (imdbn.py:42-66).  Both constructor signatures are accepted, plus
    the monolith's extra ``logging_cfg`` keyword (gdbn_model_complete.py:596)."""

    WARMUP_Y_EPOCHS = 8          # imdbn.py:540
    N_CANDIDATES = 5             # Kbuf, imdbn.py:452 (BASELINE config 4 generalises it to 64)

    def __init__(self, layer_sizes_img: list, layer_sizes_txt_or_joint=None,
                 joint_layer_size: Optional[int] = None, params: Optional[dict] = None,
                 dataloader=None, val_loader=None, device=None, text_posenc_dim: int = 0,
                 num_labels: int = 32, embedding_dim: int = 64, wandb_run=None,
                 logging_config_path: Optional[str] = None, logging_cfg: Optional[dict] = None):
        super().__init__()
        if isinstance(layer_sizes_txt_or_joint, (list, tuple)):
            if joint_layer_size is None:
                raise ValueError("joint_layer_size required with legacy constructor signature")
        elif joint_layer_size is None:
            joint_layer_size = int(layer_sizes_txt_or_joint)

        self.params = params or {}
        self.device = torch.device(device) if device is not None else torch.device(
            "cuda" if torch.cuda.is_available() else "cpu")
        self.dataloader = dataloader
        self.val_loader = val_loader
        self.wandb_run = wandb_run
        self.logging_cfg = dict(logging_cfg or {})
        self.num_labels = int(num_labels)

        try:                                                     # imdbn.py:137-145
            vb_imgs, vb_lbls = next(iter(val_loader))
            self.validation_images = vb_imgs[:8].to(self.device)
            self.validation_labels = vb_lbls[:8].to(self.device)
            self.val_batch = (vb_imgs, vb_lbls)
        except Exception:
            self.validation_images = None
            self.validation_labels = None
            self.val_batch = None

        self.image_idbn = iDBN(layer_sizes=layer_sizes_img, params=self.params,
                               dataloader=self.dataloader, val_loader=self.val_loader,
                               device=self.device, wandb_run=self.wandb_run,
                               logging_config_path=logging_config_path)
        self.features = self.image_idbn.features
        self.Dz_img = int(self.image_idbn.layers[-1].num_hidden)
        self._build_joint(Dz_img=self.Dz_img, joint_hidden=joint_layer_size)

        self.joint_cd = int(self.params.get("JOINT_CD", self.params.get("CD", 1)))     # imdbn.py:164-167
        self.cross_steps = int(self.params.get("CROSS_GIBBS_STEPS", 50))
        self.aux_every_k = int(self.params.get("JOINT_AUX_EVERY_K", 0))
        self.aux_cond_steps = int(self.params.get("JOINT_AUX_COND_STEPS", 50))
        self.arch_str = f"IMG{'-'.join(map(str, layer_sizes_img))}_JOINT{joint_layer_size}"
        self.metrics_history = []

    def _build_joint(self, Dz_img: int, joint_hidden: int):
        """Joint RBM over [z (Dz) | y (K, one softmax group)]  (imdbn.py:191-214)."""
        self.Dz_img = int(Dz_img)
        K = self.num_labels
        p = self.params
        self.joint_rbm = RBM(
            num_visible=self.Dz_img + K, num_hidden=int(joint_hidden),
            learning_rate=p.get("JOINT_LEARNING_RATE", p.get("LEARNING_RATE", 0.1)),
            weight_decay=p.get("WEIGHT_PENALTY", 0.0001), momentum=p.get("INIT_MOMENTUM", 0.5),
            dynamic_lr=p.get("LEARNING_RATE_DYNAMIC", True),
            final_momentum=p.get("FINAL_MOMENTUM", 0.95),
            softmax_groups=[(self.Dz_img, self.Dz_img + K)],
        ).to(self.device)

    # ------------------------------------------------------------------ bias init
    @torch.no_grad()
    def init_joint_bias_from_data(self, n_batches: int = 10):
        """Visible biases of the joint RBM from data statistics + per-class latent means
        (imdbn.py:216-292).  One pass over <= n_batches batches; the per-class sums come from one
        kernel (``imdbn_class_stats``) instead of K masked reductions with a host sync each."""
        if not hasattr(self, "Dz_img"):
            self.Dz_img = int(self.joint_rbm.num_visible) - self.num_labels
        Dz, K, dev = self.Dz_img, self.num_labels, self.device
        sum_z = torch.zeros(Dz, device=dev)
        class_sum = torch.zeros(K, Dz, device=dev)
        class_count = torch.zeros(K, device=dev)
        label_sum = torch.zeros(K, device=dev)
        n = 0
        for b, (imgs, lbls) in enumerate(self.dataloader):
            if b >= n_batches:
                break
            z = self.image_idbn.represent(imgs)
            y = L.f32c(lbls, z.device)
            ctx, st = L.context_for(z)
            ctx.check(ctx.lib.imdbn_class_stats(ctx.handle, L.ptr(z), L.ptr(y), z.shape[0], Dz, K,
                                                L.ptr(sum_z), L.ptr(class_sum), L.ptr(class_count),
                                                L.ptr(label_sum), st), "imdbn_class_stats")
            n += z.shape[0]
        if n == 0:
            return
        mean_z = (sum_z / n).clamp(1e-4, 1 - 1e-4)
        priors = label_sum / label_sum.sum().clamp(min=1)
        priors = (priors + 1e-6) / (priors.sum() + 1e-6 * K)
        present = (class_count > 0).unsqueeze(1)
        self.z_class_mean = torch.where(present, class_sum / class_count.clamp(min=1).unsqueeze(1),
                                        mean_z.unsqueeze(0).expand(K, Dz)).contiguous()
        self.z_class_count = class_count
        self.joint_rbm.vis_bias.data[:Dz] = torch.log(mean_z) - torch.log1p(-mean_z)
        self.joint_rbm.vis_bias.data[Dz:Dz + K] = torch.log(priors)

    # ------------------------------------------------------------------ loaders
    def load_pretrained_image_idbn(self, path: str) -> bool:
        """imdbn.py:294-342: accepts a ``{"layers": ...}`` dict or an object with ``.layers``."""
        try:
            with open(path, "rb") as f:
                obj = pickle.load(f)
        except Exception as e:
            print(f"[load_pretrained_image_idbn] error: {e}")
            return False
        if isinstance(obj, dict) and "layers" in obj:
            self.image_idbn.layers = obj["layers"]
        elif hasattr(obj, "layers"):
            self.image_idbn = obj
            if not hasattr(self.image_idbn, "text_flag"):
                self.image_idbn.text_flag = False
            if not hasattr(self.image_idbn, "arch_dir"):
                self.image_idbn.arch_dir = os.path.join("logs-idbn", "loaded")
                os.makedirs(self.image_idbn.arch_dir, exist_ok=True)
        else:
            print("[load_pretrained_image_idbn] unrecognized format")
            return False
        for rbm in self.image_idbn.layers:
            rbm.to(self.device)
            rbm.W_m = torch.zeros_like(rbm.W.data)
            rbm.hb_m = torch.zeros_like(rbm.hid_bias.data)
            rbm.vb_m = torch.zeros_like(rbm.vis_bias.data)
            if not hasattr(rbm, "softmax_groups"):
                rbm.softmax_groups = []
        dz_pre = int(self.image_idbn.layers[-1].num_hidden)
        if dz_pre != getattr(self, "Dz_img", dz_pre):
            print(f"[load_pretrained_image_idbn] rebuilding joint: Dz_img -> {dz_pre}")
            self._build_joint(Dz_img=dz_pre, joint_hidden=self.joint_rbm.num_hidden)
        print(f"[load_pretrained_image_idbn] loaded from {path}")
        return True

    def finetune_image_last_layer(self, epochs: int = 0, lr_scale: float = 0.3,
                                  cd_k: Optional[int] = None):
        """CD on the top image layer with a scaled learning rate (imdbn.py:344-384)."""
        if epochs <= 0:
            return
        last = self.image_idbn.layers[-1]
        old_lr = float(last.lr)
        last.lr = max(1e-8, old_lr * float(lr_scale))
        use_cd = int(cd_k) if cd_k is not None else int(self.image_idbn.cd_k)
        print(f"[finetune_image_last_layer] epochs={epochs}, lr={last.lr:.4g}, CD={use_cd}")
        for ep in range(int(epochs)):
            losses = []
            for batch in prefetch_to_device(self.dataloader, self.device):
                v = _flat(batch[0], self.device)
                for rbm in self.image_idbn.layers[:-1]:
                    v = rbm.forward(v)
                losses.append(last.train_epoch(v, ep, epochs, CD=use_cd))
            if self.wandb_run and losses:
                self.wandb_run.log({"img_last/finetune_loss": float(torch.stack(losses).mean()),
                                    "epoch_ft": ep})
        last.lr = old_lr
        print("[finetune_image_last_layer] done")

    # ------------------------------------------------------------------ cross-modal inference
    def _block_mask(self, B: int, known_img: bool):
        Dz, K = self.Dz_img, self.num_labels
        km = torch.zeros(B, Dz + K, device=self.device)
        if known_img:
            km[:, :Dz] = 1.0
        else:
            km[:, Dz:] = 1.0
        return km

    @torch.no_grad()
    def _cross_reconstruct(self, z_img: torch.Tensor, y_onehot: torch.Tensor,
                           steps: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """IMG->TXT conditional Gibbs, then TXT->IMG noisy mean-field annealing with mu-pull and
        best-of-K (K = ``self.N_CANDIDATES``) picked by free energy when ``joint_rbm`` has a
        ``free_energy`` method (imdbn.py:386-488).  Returns (img_from_txt [B,D], p_y_given_img [B,K])."""
        if steps is None:
            steps = self.cross_steps
        jr = self.joint_rbm
        z_img = L.f32c(z_img, self.device)
        y_onehot = L.f32c(y_onehot, self.device)
        B, Dz, K = z_img.size(0), self.Dz_img, self.num_labels
        V = Dz + K

        v_known = torch.zeros(B, V, device=self.device)
        v_known[:, :Dz] = z_img
        v_i2t = jr.conditional_gibbs(v_known, self._block_mask(B, True), n_steps=steps,
                                     sample_h=False, sample_v=False, clamp_prefix=Dz)
        p_y_given_img = v_i2t[:, Dz:]

        v_known = torch.zeros(B, V, device=self.device)
        v_known[:, Dz:] = y_onehot
        km = self._block_mask(B, False)
        if getattr(self, "z_class_mean", None) is not None:
            mu_k = self.z_class_mean[y_onehot.argmax(dim=1)]
            jr._mu_pull = {"mu_k": mu_k, "eta0": 0.15}
        else:
            jr._mu_pull = None
        v_chain = jr.noisy_meanfield_annealed(v_known=v_known, known_mask=km, n_steps=steps,
                                              T0=3.0, T1=1.0, sigma0=0.9, hot_frac=0.7,
                                              sharpen_last=3, T_cold_plus=0.9, clamp_suffix=Dz)
        n_ref = int(self.N_CANDIDATES) - 1
        if hasattr(jr, "free_energy"):
            cands = [v_chain]
            for _ in range(n_ref):
                cands.append(jr.noisy_meanfield_annealed(
                    v_known=cands[-1], known_mask=km, n_steps=1, T0=0.9, T1=0.9, sigma0=0.0,
                    hot_frac=0.0, sharpen_last=0, T_cold_plus=0.9, clamp_suffix=Dz))
            stack = torch.stack(cands, dim=0).contiguous()
            Fe = torch.stack([jr.free_energy(c) for c in cands], dim=0).contiguous()
            v_pick = torch.empty_like(v_chain)
            ctx, st = L.context_for(v_pick)
            ctx.check(ctx.lib.imdbn_best_of_k(ctx.handle, L.ptr(stack), L.ptr(Fe), len(cands), B, V,
                                              L.ptr(v_pick), None, st), "imdbn_best_of_k")
        else:
            # Without the hook every candidate scores 0 and arg-min returns the main chain
            # (imdbn.py:454-474): the refinements cannot influence the result, so only
            # their random-field call numbers are consumed.
            jr._rng_stream += n_ref
            v_pick = v_chain
        jr._mu_pull = None

        z_from_y = v_pick[:, :Dz]
        if hasattr(self, "z_affine_scale") and hasattr(self, "z_affine_bias"):
            z_from_y = (z_from_y - self.z_affine_bias) / (self.z_affine_scale + 1e-6)
        return self.image_idbn.decode(z_from_y.contiguous()), p_y_given_img

    @torch.no_grad()
    def represent(self, batch: Tuple[torch.Tensor, torch.Tensor]) -> torch.Tensor:
        """Joint hidden probabilities of (images, labels)  (imdbn.py:490-506)."""
        img, lbl = batch
        z = self.image_idbn.represent(img)
        y = L.f32c(lbl, z.device)
        return self.joint_rbm.forward(torch.cat([z, y], dim=1))

    # ------------------------------------------------------------------ joint training
    def train_joint(self, epochs: int, log_every_pca: int = 25, log_every_probe: int = 10,
                    log_every: int = 5, w_rec: float = 1.0, w_sup: float = 0.0):
        """Warm-up (label-clamped aux CD) then free CD + aux clamped CD, with the cross-modal metrics
        of every batch (imdbn.py:508-639).  ``metrics_history`` gets one dict per epoch."""
        print("[iMDBN] joint training (with warmup y-clamp)")
        self.init_joint_bias_from_data(n_batches=10)
        jr, dev = self.joint_rbm, self.device
        Dz, K = self.Dz_img, self.num_labels
        aux_steps = int(self.params.get("JOINT_AUX_COND_STEPS", 10))          # imdbn.py:564
        clamp_kw = dict(CD=1, cond_init_steps=aux_steps, sample_h=False, sample_v=False,
                        aux_lr_mult=0.3, use_noisy_init=True)

        for epoch in range(int(epochs)):
            cd_losses = []
            acc = torch.zeros(4, device=dev, dtype=torch.float64)   # top1, top3, ce, mse sums
            n_seen, npix = 0, None
            for b_idx, (img, y) in enumerate(prefetch_to_device(self.dataloader, dev)):
                img = _flat(img, dev)
                y = L.f32c(y, dev)
                z_img = self.image_idbn.represent(img)
                B = z_img.size(0)
                vk_y = torch.zeros(B, Dz + K, device=dev)
                vk_y[:, Dz:] = y
                km_y = self._block_mask(B, False)
                if epoch < self.WARMUP_Y_EPOCHS:
                    for _ in range(2):
                        jr.train_epoch_clamped(vk_y, km_y, epoch, epochs, **clamp_kw)
                else:
                    cd_losses.append(jr.train_epoch(torch.cat([z_img, y], dim=1), epoch, epochs,
                                                    CD=self.joint_cd))
                    jr.train_epoch_clamped(vk_y, km_y, epoch, epochs, reclamp_negative=False, **clamp_kw)
                    if (b_idx % 50) == 0:
                        vk_z = torch.zeros(B, Dz + K, device=dev)
                        vk_z[:, :Dz] = z_img
                        jr.train_epoch_clamped(vk_z, self._block_mask(B, True), epoch, epochs,
                                               reclamp_negative=False, **clamp_kw)

                img_from_txt, p_y = self._cross_reconstruct(z_img, y, steps=self.cross_steps)
                gt = y.argmax(dim=1)
                topk = p_y.topk(k=min(3, p_y.size(1)), dim=1).indices
                ce = F.binary_cross_entropy(p_y.clamp(1e-6, 1 - 1e-6),
                                            F.one_hot(gt, num_classes=p_y.size(1)).float(),
                                            reduction="sum")
                mse = F.mse_loss(img_from_txt.view_as(img), img, reduction="sum")
                acc += torch.stack([(p_y.argmax(dim=1) == gt).sum().double(),
                                    (topk == gt.unsqueeze(1)).any(dim=1).sum().double(),
                                    ce.double(), mse.double()])
                n_seen += B
                npix = img.size(1)

            if n_seen > 0:
                top1, top3, ce_sum, mse_sum = (float(x) for x in acc.cpu())
                rec = {"epoch": epoch,
                       "cross_modality/text_top1": top1 / n_seen,
                       "cross_modality/text_top3": top3 / n_seen,
                       "cross_modality/text_ce": ce_sum / n_seen,
                       "cross_modality/image_mse": mse_sum / max(1, n_seen * max(1, npix or 1))}
                if cd_losses:
                    rec["joint/cd_loss"] = float(torch.stack(cd_losses).mean())
                self.metrics_history.append(rec)
                if self.wandb_run:
                    self.wandb_run.log(rec)
            # linear probes on the joint embeddings (imdbn.py:695-706), on the device
            if (self.wandb_run and self.val_loader is not None and self.features is not None
                    and log_every_probe and epoch % log_every_probe == 0):
                from .probe_utils import log_joint_linear_probe
                try:
                    log_joint_linear_probe(self, epoch=epoch, n_bins=5, test_size=0.2, steps=1000, lr=1e-2,
                                           patience=20, min_delta=0.0, metric_prefix="joint")
                except Exception as e:
                    self.wandb_run.log({"warn/joint_probe_error": str(e)})
            if epoch % max(1, int(log_every)) == 0:
                self._log_snapshots(epoch)
        print("[iMDBN] joint training finished.")

    def _log_snapshots(self, epoch: int, num: int = 8):
        """W&B image/confusion rendering (imdbn.py:714-813): out of scope for the CUDA path."""
        return None

    # ------------------------------------------------------------------ checkpoints
    def save_model(self, path: str):
        """Dual-format pickle (imdbn.py:815-883): DBN-compatible ``layers``/``params`` plus the full
        iMDBN components."""
        all_layers = list(self.image_idbn.layers) + [self.joint_rbm]
        for l in all_layers:
            l.sync_momenta()                 # no-op unless peer-memory data parallelism is active
        payload = {
            "layers": all_layers, "params": self.params,
            "image_idbn": self.image_idbn, "joint_rbm": self.joint_rbm,
            "num_labels": self.num_labels, "Dz_img": self.Dz_img, "arch_str": self.arch_str,
            "features": self.features,
            "metadata": {"saved_at": datetime.datetime.now().isoformat(), "model_type": "iMDBN",
                         "architecture": self.arch_str},
        }
        for name in ("z_class_mean", "z_affine_scale", "z_affine_bias", "class_names"):
            if getattr(self, name, None) is not None:
                payload[name] = getattr(self, name)
        with open(path, "wb") as f:
            pickle.dump(payload, f)
        print(f"[iMDBN] Model saved to {path}")
        print(f"[iMDBN] Architecture: {self.arch_str}")
        print(f"[iMDBN] Total layers: {len(all_layers)} (image: {len(self.image_idbn.layers)}, joint: 1)")

    @staticmethod
    def load_model(path: str, device=None) -> Dict[str, Any]:
        """imdbn.py:885-934: returns the payload dict with the RBMs moved to ``device``."""
        if device is None:
            device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        with open(path, "rb") as f:
            payload = pickle.load(f)
        if "image_idbn" in payload:
            for rbm in payload["image_idbn"].layers:
                rbm.to(device)
        if "joint_rbm" in payload:
            payload["joint_rbm"].to(device)
        print(f"[iMDBN] Model loaded from {path}")
        return payload
