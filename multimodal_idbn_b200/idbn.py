"""Drop-in ``iDBN`` (reference ``imdbn/models/idbn.py``): a stack of CUDA-backed RBMs.

Hot loop = ``train`` (idbn.py:195-204): per minibatch every layer takes one CD-k update and the
batch is pushed through the *updated* layer.  Here the loop issues no host synchronisation: losses
stay on the device until the end of the epoch, and the next minibatch is copied host->device on a
side stream while the current one trains.  ``represent`` / ``reconstruct`` / ``decode`` are chains
of up / down passes; ``save_model`` writes the reference's ``{"layers", "params"}`` pickle.

W&B visualisation (PCA plots: idbn.py:207-284) is out of scope; the scalar loss and the linear-probe
accuracies of the monitored layers (idbn.py:286-305, ``probe_utils.py`` here) are logged when a ``wandb_run`` is given.
"""
from __future__ import annotations

import ctypes as C
import os
import pickle
from typing import Iterable, List, Optional

import torch

from . import _lib as L
from . import dist as _dist
from .rbm import RBM


def _flat(x: torch.Tensor, device) -> torch.Tensor:
    """``img.to(device).view(B, -1).float()`` (idbn.py:200,319,336)."""
    if x.dim() == 2 and x.dtype is torch.float32 and x.device == device:      # already in place (the hot loop)
        return x
    x = x.to(device, non_blocking=True)
    return x.reshape(x.size(0), -1).float()


# device staging buffers survive the generator (a new epoch reuses them instead of asking the allocator, whose
# cudaMalloc would synchronise the device): per (device, ring) the buffers and an event recorded on the consumer's
# stream when the previous generator finished
_staging_cache = {}


def prefetch_to_device(loader: Iterable, device, ring: int = 4):
    """Yield the loader's batches already on ``device``: batch i+1 is copied on a side stream while the caller
    works on batch i.  Falls back to plain iteration on CPU devices.

    The device copies live in a ring of ``ring`` preallocated buffers per tensor slot that survives the generator
    (no allocator traffic in the steady state and none at the start of the next epoch -- a cudaMalloc would
    synchronise the device).  Contract for the consumer: a yielded tensor stays valid while the NEXT yielded batch
    is being processed (one batch of lookahead, as ``iDBN.train`` uses it); the copy into a buffer is ordered,
    by an event, after everything the consumer had enqueued on its stream when the copy was issued, and the
    buffer's previous occupant is at least ``ring - 1`` batches old by then.
    (A worker-thread variant was measured and dropped: 183 us per C2 step against 140-150 us -- the GIL hand-offs
    cost more than the ~20 us of staging work they take off the consumer thread.)"""
    device = torch.device(device)
    if device.type != "cuda" or getattr(loader, "device_resident", False):
        # (datasets.DeviceLoader: the batches are views of a matrix that already lives in HBM)
        for batch in loader:
            yield batch
        return
    ring = max(3, int(ring))
    main = torch.cuda.current_stream(device)
    cache_key = (device.index if device.index is not None else torch.cuda.current_device(), ring)
    # buffers, copy stream and events survive the generator: creating them costs ~100 us, a fifth of a 20-batch epoch
    bufs, prev_end, copy_stream, fences, dones = _staging_cache.pop(cache_key, ({}, None, None, None, None))
    if copy_stream is None:
        copy_stream = torch.cuda.Stream(device=device)
        fences = [torch.cuda.Event() for _ in range(ring)]
        dones = [torch.cuda.Event() for _ in range(ring)]
    if prev_end is not None:
        copy_stream.wait_event(prev_end)           # the previous consumer's reads of these buffers

    copy_h2d = L.load_library().imdbn_copy_async
    raw_copy_stream = C.c_void_p(copy_stream.cuda_stream)

    def stage(batch, n):
        slot = n % ring
        fence = fences[slot]
        fence.record(main)         # everything the consumer enqueued so far (it is >= 2 batches behind this one)
        copy_stream.wait_event(fence)
        moved = []
        for pos, b in enumerate(batch):
            if not torch.is_tensor(b):
                moved.append(b)
                continue
            dst = bufs.get((slot, pos))
            if dst is None or dst.shape != b.shape or dst.dtype != b.dtype:
                dst = torch.empty(b.shape, dtype=b.dtype, device=device)
                dst.record_stream(copy_stream)
                bufs[(slot, pos)] = dst
            if b.device.type == "cpu" and b.is_contiguous():
                # raw cudaMemcpyAsync on the copy stream: no stream-context switch, no dispatcher (host time per
                # step is within 10 % of device time at batch 64)
                rc = copy_h2d(dst.data_ptr(), b.data_ptr(), b.numel() * b.element_size(), raw_copy_stream)
                if rc:
                    raise RuntimeError(f"imdbn_copy_async failed (cudaError {rc})")
            else:
                with torch.cuda.stream(copy_stream):
                    dst.copy_(b, non_blocking=True)
            moved.append(dst)
        ev = dones[slot]
        ev.record(copy_stream)
        return tuple(moved), ev

    it = iter(loader)
    n = 0
    try:
        try:
            cur = stage(next(it), n)
        except StopIteration:
            return
        while cur is not None:
            n += 1
            try:
                nxt = stage(next(it), n)
            except StopIteration:
                nxt = None
            moved, ev = cur
            main.wait_event(ev)
            yield moved
            cur = nxt
    finally:
        end = prev_end if prev_end is not None else torch.cuda.Event()
        end.record(main)
        _staging_cache[cache_key] = (bufs, end, copy_stream, fences, dones)


class iDBN:
    """Image deep belief network: ``layers`` is a list of :class:`RBM` (idbn.py:39-161)."""

    def __init__(self, layer_sizes: List[int], params: dict, dataloader, val_loader, device,
                 wandb_run=None, logging_config_path: Optional[str] = None):
        self.layers: List[RBM] = []
        self.params = params
        self.dataloader = dataloader
        self.val_loader = val_loader
        self.device = torch.device(device) if device is not None else torch.device(
            "cuda" if torch.cuda.is_available() else "cpu")
        self.wandb_run = wandb_run

        self.logging_cfg = {}
        try:                                                      # idbn.py:98-110
            import yaml
            from pathlib import Path
            cfg_path = Path(logging_config_path) if logging_config_path else Path(
                "src/configs/logging_config.yaml")
            if cfg_path.exists():
                with cfg_path.open("r") as f:
                    cfg = yaml.safe_load(f)
                if isinstance(cfg, dict):
                    self.logging_cfg = cfg
        except Exception:
            pass

        self.text_flag = False
        self.arch_str = "-".join(map(str, layer_sizes))
        self.arch_dir = os.path.join("logs-idbn", f"architecture_{self.arch_str}")
        try:
            os.makedirs(self.arch_dir, exist_ok=True)             # idbn.py:115-116
        except OSError:
            pass

        self.cd_k = int(self.params.get("CD", 1))
        self.sparsity_last = bool(self.params.get("SPARSITY", False))
        self.sparsity_factor = float(self.params.get("SPARSITY_FACTOR", 0.1))

        try:
            self.val_batch, self.val_labels = next(iter(val_loader))
        except Exception:
            self.val_batch, self.val_labels = None, None

        self.features = None
        try:                                                      # idbn.py:130-146
            indices = val_loader.dataset.indices
            base = val_loader.dataset.dataset
            feats = {
                "Cumulative Area": torch.tensor([base.cumArea_list[i] for i in indices], dtype=torch.float32),
                "Convex Hull": torch.tensor([base.CH_list[i] for i in indices], dtype=torch.float32),
                "Labels": torch.tensor([base.labels[i] for i in indices], dtype=torch.float32),
            }
            dens = getattr(base, "density_list", None)
            if dens is not None:
                feats["Density"] = torch.tensor([dens[i] for i in indices], dtype=torch.float32)
            self.features = feats
        except Exception:
            pass

        n = len(layer_sizes) - 1
        for i in range(n):                                        # idbn.py:149-161
            self.layers.append(RBM(
                num_visible=layer_sizes[i], num_hidden=layer_sizes[i + 1],
                learning_rate=self.params["LEARNING_RATE"],
                weight_decay=self.params["WEIGHT_PENALTY"],
                momentum=self.params["INIT_MOMENTUM"],
                dynamic_lr=self.params["LEARNING_RATE_DYNAMIC"],
                final_momentum=self.params["FINAL_MOMENTUM"],
                sparsity=(self.sparsity_last and i == n - 1),
                sparsity_factor=self.sparsity_factor,
            ).to(self.device))
        self.loss_history: List[float] = []

    def _layers_to_monitor(self) -> List[int]:
        layers = {len(self.layers)}
        if len(self.layers) > 1:
            layers.add(1)
        return sorted(layers)

    def _layer_tag(self, idx: int) -> str:
        return f"layer{idx}"

    # ------------------------------------------------------------------ training
    @torch.no_grad()
    def train_step(self, v: torch.Tensor, epoch: int, epochs: int, next_v: Optional[torch.Tensor] = None,
                   loss_out: Optional[torch.Tensor] = None) -> List[torch.Tensor]:
        """One minibatch of the hot loop (idbn.py:200-204); returns the per-layer losses as device
        scalars (no host synchronisation).  ``next_v`` = the next minibatch if already known (flattened
        fp32 on the device, the SAME tensor object that will be passed as ``v`` next time): the first
        layer then computes its positive phase in this step's post-update forward pass.  ``loss_out`` =
        optional fp32 vector with one element per layer (device or PINNED host memory) that the kernels
        write the losses into directly; the returned list then holds views of it."""
        v = _flat(v, self.device)
        if v.device.type == "cuda" and _dist.state() is None and getattr(self, "fused_step", True):
            return self._train_step_fused(v, epoch, epochs, next_v, loss_out)
        losses = []
        for i, rbm in enumerate(self.layers):
            loss, v = rbm.train_epoch_fwd(v, epoch, epochs, CD=self.cd_k, next_data=next_v if i == 0 else None,
                                          loss_out=None if loss_out is None else loss_out[i])
            losses.append(loss)
        return losses

    def _train_step_fused(self, v, epoch, epochs, next_v, loss_out) -> List[torch.Tensor]:
        """All layers of one minibatch through ONE library call (``imdbn_idbn_train_step``): the host pays for
        one foreign call instead of one per layer plus the event / stream bookkeeping.

        ``pipeline_layers = True``: layer 0 of minibatch t+1 does not depend on the upper layers of minibatch t
        (they only consume layer 0's forward output), so the two run CONCURRENTLY on disjoint SM partitions
        (CUDA green contexts, ``imdbn_sm_partition``: ``pipeline_reserve_sms`` SMs for the upper layers, the rest
        for layer 0; every persistent grid is sized to its partition).  Without partitions (driver too old,
        ``pipeline_partition = False``) the upper layers go to an ordinary side stream, layer 0's grids are capped
        at ``num_sms - pipeline_reserve_sms`` and launched without early dependent launch (an early-launched grid
        parks its CTAs on exactly the SMs that were left free).  Measured on B200, C2: 141-150 us per step with
        partitions, 165 us with the fallback, 174 us unpipelined.  The arithmetic is the same; the split-K
        summation order follows the grid sizes, so results agree with the unpipelined run to rounding.
        ``sync()`` orders readers of the losses (which then live in a ring reused every 8 steps) and of the
        parameters on the caller's stream."""
        dev = v.device
        n = len(self.layers)
        B = v.shape[0]
        if not v.is_contiguous():
            v = v.contiguous()
        nxt = _flat(next_v, dev) if next_v is not None else None
        if nxt is not None and not nxt.is_contiguous():
            nxt = nxt.contiguous()
        Bn = nxt.shape[0] if nxt is not None else 0
        d = self.__dict__
        piped = bool(d.get("pipeline_layers", False)) and n > 1
        reserve = int(d.get("pipeline_reserve_sms", 16)) if piped else 0
        s_caller = torch.cuda.current_stream(dev).cuda_stream
        st = d.get("_fused")
        key = (B, Bn, dev, s_caller, piped, reserve, n, id(self.layers[-1]))
        if st is None or st["key"] != key:
            if st is not None:
                # the old rings / loss ring may still be read and written by the previous steps' kernels on the side
                # or partition streams: join them into the caller's stream, on which the caching allocator orders the
                # re-use of the storage we are about to drop
                self.sync()
                for s_old in st.get("streams") or ():
                    for ring in st["rings"]:
                        for t in ring:
                            t.record_stream(s_old)
                    if st.get("loss_ring") is not None:
                        st["loss_ring"].record_stream(s_old)
            ctx_c, s_caller = L.context_for(v)
            rings = [[torch.empty(B + (Bn if l == 0 else 0), r.num_hidden, device=dev, dtype=torch.float32)
                      for l, r in enumerate(self.layers)] for _ in range(2)]
            st = dict(key=key, rings=rings, parity=0, ctx0=ctx_c, s0=s_caller, ctx1=None, s1=0, early=1, streams=[],
                      fwd=[(C.c_void_p * n)(*[t.data_ptr() for t in ring]) for ring in rings],
                      rbms=(L.RbmStruct * n)(), upds=(L.UpdateStruct * n)(), rngs=(L.RngStruct * n)(),
                      loss=(C.c_void_p * n)())
            n_sms = torch.cuda.get_device_properties(dev).multi_processor_count
            idx = dev.index if dev.index is not None else torch.cuda.current_device()
            part = L.sm_partition(idx, reserve) if piped and reserve >= 8 and getattr(self, "pipeline_partition", True) else None
            if part is not None:
                # layer 0 and the upper layers on two streams with DISJOINT SMs (green contexts): early dependent
                # launch stays on, every grid is sized to its own partition
                big, small, n_big, n_small = part
                st["ctx0"], st["s0"] = L.context_for_stream(idx, big), big
                st["ctx1"], st["s1"] = L.context_for_stream(idx, small), small
                st["ctx0"].set_sm_limit(n_big)
                st["ctx1"].set_sm_limit(n_small)
                st["streams"] = [torch.cuda.ExternalStream(big, device=dev), torch.cuda.ExternalStream(small, device=dev)]
            elif piped:
                # PRIVATE contexts (not the shared per-stream ones): an SM limit set on a shared context would leak
                # into every other call on that stream
                side = torch.cuda.Stream(device=dev)
                st["ctx0"] = L.private_context(idx)
                st["ctx1"], st["s1"] = L.private_context(idx), side.cuda_stream
                st["streams"] = [side]
                st["early"] = 0
                st["ctx0"].set_sm_limit(max(8, n_sms - reserve) if reserve > 0 else 0)
                if reserve > 0:              # the side stream's persistent grids stay inside the SMs left to them
                    st["ctx1"].set_sm_limit(reserve)
            self._fused = st
            self._side_stream = st["streams"]
        par = st["parity"]
        st["parity"] = par ^ 1
        # losses for the host: the kernels never store to host memory themselves when the layers are pipelined -- the
        # system-scope flush at the end of the (critical-path) kernel that wrote the loss queues behind the minibatch
        # copy on PCIe (measured: +10 us per C2 step); they write the device ring and ONE small copy on the upper layers'
        # stream, which has slack, carries the losses of the step to the caller's buffer
        host_loss = loss_out is not None and piped and not loss_out.is_cuda
        if host_loss:
            if loss_out.dtype != torch.float32 or loss_out.numel() != n or not loss_out.is_contiguous():
                raise ValueError("loss_out must be a contiguous fp32 vector with one element per layer")
        if (loss_out is None or host_loss) and piped:
            # the upper layers write their losses from the side stream: use model-owned storage (a ring reused
            # every 8 steps) instead of a fresh allocation the caching allocator might recycle too early
            ring = st.get("loss_ring")
            if ring is None:
                ring = st["loss_ring"] = torch.zeros(8, n, device=dev, dtype=torch.float32)
            st["loss_pos"] = (st.get("loss_pos", -1) + 1) % 8
            loss_t = ring[st["loss_pos"]]
        elif loss_out is None:
            loss_t = torch.empty(n, device=dev, dtype=torch.float32)
        else:
            store = loss_out.untyped_storage().data_ptr()
            if st.get("loss_store_ok") != store:           # (is_pinned() is a driver query: once per storage)
                if loss_out.dtype != torch.float32 or not (loss_out.is_cuda or loss_out.is_pinned()):
                    raise ValueError("loss_out must be fp32 on the device or in pinned host memory")
                st["loss_store_ok"] = store
            if loss_out.numel() != n:
                raise ValueError("loss_out must hold one element per layer")
            loss_t = loss_out
        base = loss_t.data_ptr()
        st["ctx0"].set_precision(L.current_precision())
        ident = st.setdefault("ident", [None] * n)
        rngs, loss_ptrs = st["rngs"], st["loss"]
        for l, rbm in enumerate(self.layers):
            # the argument structs are rebuilt only when something they describe changed (parameter storage,
            # epoch-dependent hyper-parameters); per step only the call number of the random field moves
            rd = rbm.__dict__
            wm = rd.get("W_m")
            hbm, vbm, ps = rd.get("hb_m"), rd.get("vb_m"), rbm._parameters
            key = (ps["W"].data_ptr(), ps["hid_bias"].data_ptr(), ps["vis_bias"].data_ptr(),
                   wm.data_ptr() if wm is not None else 0, hbm.data_ptr() if hbm is not None else 0,
                   vbm.data_ptr() if vbm is not None else 0, epoch, rd["lr"], rd["sparsity"], rd["momentum"],
                   rd["final_momentum"], rd["weight_decay"], rd["sparsity_factor"], rd["dynamic_lr"],
                   tuple(map(tuple, rd.get("softmax_groups") or ())))
            if ident[l] != key:
                lr, mom = rbm._hyper(epoch)
                st["rbms"][l] = rbm._struct(training=True)          # (may re-home the momenta: rebuild the key after)
                st["upds"][l] = rbm._update_struct(lr, mom, B, rbm.sparsity)
                ident[l] = (rbm.W.data_ptr(), rbm.hid_bias.data_ptr(), rbm.vis_bias.data_ptr(), rbm.W_m.data_ptr(),
                            rbm.hb_m.data_ptr(), rbm.vb_m.data_ptr(), epoch, rbm.lr, rbm.sparsity, rbm.momentum,
                            rbm.final_momentum, rbm.weight_decay, rbm.sparsity_factor, rbm.dynamic_lr,
                            tuple(map(tuple, rbm.softmax_groups or ())))
            if "_rng_seed" not in rd:                    # object un-pickled from a reference checkpoint
                rngs[l] = rbm._next_rng()
            else:
                r = rngs[l]
                r.seed, r.stream, r.row0 = rd["_rng_seed"], rd["_rng_stream"] & 0xFFFFFFFF, 0
                rd["_rng_stream"] += 1
            loss_ptrs[l] = base + 4 * l
        first = self.layers[0]
        cached = first.__dict__.pop("_pos_cache", None)
        pos_in = cached[1] if cached is not None and cached[0] == first._pos_key(v) else None
        ctx0, ctx1 = st["ctx0"], st["ctx1"]
        ctx0.check(ctx0.lib.imdbn_idbn_train_step(
            ctx0.handle, ctx1.handle if ctx1 is not None else None, n, st["rbms"], st["upds"], st["rngs"],
            L.ptr(v), B, int(self.cd_k), L.ptr(pos_in), L.ptr(nxt), Bn, st["fwd"][par], st["loss"], par, st["s0"],
            st["s1"], s_caller, st["early"]), "imdbn_idbn_train_step")
        if host_loss:
            ctx0.check(ctx0.lib.imdbn_copy_async(loss_out.data_ptr(), base, 4 * n, st["s1"]), "imdbn_copy_async")
        for rbm in self.layers:
            rd = rbm.__dict__
            rd["_n_updates"] = rd.get("_n_updates", 0) + 1
            rd.pop("_pos_cache", None)
        if nxt is not None:
            first.__dict__["_pos_cache"] = (first._pos_key(nxt), st["rings"][par][0][B:])
        return list((loss_out if host_loss else loss_t).unbind(0))

    def sync(self) -> None:
        """Make the current stream wait for the upper layers' side stream (call before reading losses or
        the upper layers' parameters on the current stream)."""
        for side in self.__dict__.get("_side_stream") or ():
            torch.cuda.current_stream(self.device).wait_stream(side)

    def train(self, epochs: int, log_every_pca: int = 25, log_every_probe: int = 10):
        """Layer-interleaved CD training (idbn.py:179-305).  ``loss_history`` receives the mean
        loss of every epoch (one device->host read per epoch)."""
        # train() owns the loop and joins the streams at the end of every epoch, so it pipelines the layers unless
        # told otherwise (pipeline_layers = False); direct train_step callers opt in explicitly
        auto = getattr(self, "pipeline_layers", None) is None
        if auto:
            self.pipeline_layers = (self.device.type == "cuda" and len(self.layers) > 1 and _dist.state() is None)
        try:
            self._train_epochs(epochs, log_every_probe)
        finally:
            self.sync()
            if auto:
                self.pipeline_layers = None

    def _train_epochs(self, epochs: int, log_every_probe: int = 10) -> None:
        def keep(step_losses):
            # pipelined layers: the losses live in a short ring written from the side stream -- copy them there
            sides = self.__dict__.get("_side_stream")
            if not sides or not getattr(self, "pipeline_layers", False):
                return step_losses
            with torch.cuda.stream(sides[-1]):
                return [torch.stack(step_losses)]

        for epoch in range(int(epochs)):
            losses: List[torch.Tensor] = []
            cur = None
            for batch in prefetch_to_device(self.dataloader, self.device):
                nxt = _flat(batch[0], self.device)
                if cur is not None:
                    losses.extend(keep(self.train_step(cur, epoch, epochs, next_v=nxt)))
                cur = nxt
            if cur is not None:
                losses.extend(keep(self.train_step(cur, epoch, epochs)))
            self.sync()
            if losses:
                mean_loss = float(torch.cat([l.reshape(-1) for l in losses]).mean())
                self.loss_history.append(mean_loss)
                if self.wandb_run:
                    self.wandb_run.log({"idbn/loss": mean_loss, "epoch": epoch})
            # linear probes on the monitored layers (idbn.py:286-305): embeddings, binning and the probe itself run on
            # the device (probe_utils.py here); the PCA figures of idbn.py:244-284 are rendering and not produced
            if (self.wandb_run and self.val_loader is not None and self.features is not None
                    and log_every_probe and epoch % log_every_probe == 0):
                from .probe_utils import log_linear_probe
                for layer_idx in self._layers_to_monitor():
                    tag = self._layer_tag(layer_idx)
                    try:
                        log_linear_probe(self, epoch=epoch, n_bins=5, test_size=0.2, steps=1000, lr=1e-2, patience=20,
                                         min_delta=0.0, upto_layer=layer_idx, layer_tag=tag)
                    except Exception as e:       # (the reference logs and continues)
                        self.wandb_run.log({f"warn/idbn_probe_error_{tag}": str(e)})

    # ------------------------------------------------------------------ inference chains
    @torch.no_grad()
    def represent(self, x: torch.Tensor, upto_layer: Optional[int] = None) -> torch.Tensor:
        """idbn.py:307-323."""
        self.sync()
        v = _flat(x, self.device)
        L = len(self.layers) if upto_layer is None else max(0, min(len(self.layers), int(upto_layer)))
        for i in range(L):
            v = self.layers[i].forward(v)
        return v

    @torch.no_grad()
    def reconstruct(self, x: torch.Tensor) -> torch.Tensor:
        """idbn.py:325-344."""
        return self.decode(self.represent(x))

    @torch.no_grad()
    def decode(self, top: torch.Tensor) -> torch.Tensor:
        """idbn.py:346-359."""
        self.sync()
        cur = top.to(self.device)
        for rbm in reversed(self.layers):
            cur = rbm.backward(cur)
        return cur

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("_fused", None)            # library handles, ring buffers and the side stream are not model state
        state.pop("_side_stream", None)
        return state

    def save_model(self, path: str):
        """``{"layers": [...], "params": ...}`` pickle (idbn.py:361-373).  With peer-memory data parallelism
        active this is a collective (every rank calls it; momenta slabs are gathered first)."""
        self.sync()
        for l in self.layers:
            l.sync_momenta()
        with open(path, "wb") as f:
            pickle.dump({"layers": self.layers, "params": self.params}, f)
        print(f"[iDBN] Model saved to {path}")
