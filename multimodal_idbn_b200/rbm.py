"""Drop-in ``RBM`` (API of the reference's ``imdbn/models/rbm.py``) whose arithmetic runs in the
sm_100a kernels of ``libimdbn_b200.so``.

Same constructor, attributes (``W, hid_bias, vis_bias`` Parameters; ``W_m, hb_m, vb_m`` plain
tensors; ``lr, weight_decay, momentum, final_momentum, dynamic_lr, sparsity, sparsity_factor,
softmax_groups``), method names, defaults and return types as the reference class
(rbm.py:24-483).  Differences, all deliberate:

* CUDA only.  A method called with the parameters on the CPU raises ``RuntimeError``.
* Random numbers come from a counter-based Philox field (``oracle/philox.py`` is the normative
  statement) addressed by ``(seed, call number, draw, row, column)`` instead of the global torch
  generator; ``set_rng(seed, stream)`` pins it.  Every stochastic method call advances the call
  number by one.
* Methods never build an autograd graph (nothing in the reference back-propagates through an RBM).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib as L
from . import dist as _dist

_instances = 0


def _schedule(t: int, t_max: int, start: float, end: float) -> float:
    if t_max <= 1:
        return float(end)
    frac = min(max(t / (t_max - 1), 0.0), 1.0)
    return float(start + (end - start) * frac)


class RBM(nn.Module):
    """Bernoulli RBM with optional softmax groups on the visible layer (reference rbm.py:24-79)."""

    def __init__(self, num_visible: int, num_hidden: int, learning_rate: float,
                 weight_decay: float, momentum: float, dynamic_lr: bool = False,
                 final_momentum: float = 0.97, sparsity: bool = False,
                 sparsity_factor: float = 0.05,
                 softmax_groups: Optional[List[Tuple[int, int]]] = None):
        super().__init__()
        self.num_visible = int(num_visible)
        self.num_hidden = int(num_hidden)
        self.lr = float(learning_rate)
        self.weight_decay = float(weight_decay)
        self.momentum = float(momentum)
        self.dynamic_lr = bool(dynamic_lr)
        self.final_momentum = float(final_momentum)
        self.sparsity = bool(sparsity)
        self.sparsity_factor = float(sparsity_factor)
        self.softmax_groups = softmax_groups or []

        dev = torch.device("cuda" if torch.cuda.is_available() else "cpu")   # rbm.py:69
        scale = math.sqrt(max(1, self.num_visible))
        self.W = nn.Parameter(torch.randn(self.num_visible, self.num_hidden, device=dev) / scale)
        self.hid_bias = nn.Parameter(torch.zeros(self.num_hidden, device=dev))
        self.vis_bias = nn.Parameter(torch.zeros(self.num_visible, device=dev))
        self.W_m = torch.zeros_like(self.W)
        self.hb_m = torch.zeros_like(self.hid_bias)
        self.vb_m = torch.zeros_like(self.vis_bias)

        global _instances
        _instances += 1
        self._rng_seed = (torch.initial_seed() + 0x9E3779B97F4A7C15 * _instances) & (2 ** 64 - 1)
        self._rng_stream = 0

    # ------------------------------------------------------------------ plumbing
    def set_rng(self, seed: int, stream: int = 0) -> None:
        """Pin the random field: the next stochastic call uses ``(seed, stream)``."""
        self._rng_seed = int(seed) & (2 ** 64 - 1)
        self._rng_stream = int(stream)

    def _next_rng(self, row0: int = 0) -> L.RngStruct:
        if not hasattr(self, "_rng_seed"):          # object un-pickled from a reference checkpoint
            self._rng_seed, self._rng_stream = torch.initial_seed() & (2 ** 64 - 1), 0
        r = L.RngStruct(self._rng_seed, self._rng_stream & 0xFFFFFFFF, row0)
        self._rng_stream += 1
        return r

    def _sync_buffers(self) -> None:
        """Momenta follow the parameters' device (the reference re-creates them by hand,
        imdbn.py:325-331); parameters must be contiguous fp32."""
        dev = self.W.device
        for name, like in (("W_m", self.W), ("hb_m", self.hid_bias), ("vb_m", self.vis_bias)):
            buf = getattr(self, name, None)
            if buf is None or buf.shape != like.shape:
                setattr(self, name, torch.zeros_like(like.data))
            elif buf.device != dev or buf.dtype != torch.float32 or not buf.is_contiguous():
                setattr(self, name, buf.to(device=dev, dtype=torch.float32).contiguous())
        for name in ("W", "hid_bias", "vis_bias"):
            p = getattr(self, name)
            if p.dtype != torch.float32 or not p.is_contiguous():
                p.data = p.data.float().contiguous()

    def _struct(self, training: bool = False) -> L.RbmStruct:
        groups = list(getattr(self, "softmax_groups", None) or [])
        if len(groups) > L.MAX_GROUPS:
            raise ValueError(f"at most {L.MAX_GROUPS} softmax groups are supported")
        if training:
            self._sync_buffers()
        s = L.RbmStruct()
        s.W, s.hb, s.vb = self.W.data_ptr(), self.hid_bias.data_ptr(), self.vis_bias.data_ptr()
        if training:
            s.Wm, s.hbm, s.vbm = self.W_m.data_ptr(), self.hb_m.data_ptr(), self.vb_m.data_ptr()
        s.V, s.H, s.ngroups = self.num_visible, self.num_hidden, len(groups)
        for i, (a, b) in enumerate(groups):
            s.group_start[i], s.group_end[i] = int(a), int(b)
        return s

    def _ctx(self):
        return L.context_for(self.W)

    def _in(self, x: torch.Tensor, width: int) -> torch.Tensor:
        x = L.f32c(x, self.W.device)
        if x.dim() != 2 or x.shape[1] != width:
            raise RuntimeError(f"expected a [B, {width}] tensor, got {tuple(x.shape)}")
        return x

    def _hyper(self, epoch: int, lr_mult: float = 1.0):
        lr = self.lr / (1 + 0.01 * epoch) if self.dynamic_lr else self.lr     # rbm.py:194
        mom = self.momentum if epoch <= 5 else self.final_momentum              # rbm.py:195
        return lr_mult * lr, mom

    # ------------------------------------------------------------------ passes
    def _up(self, v, T=1.0, sample=False, rng=None, draw=0):
        v = self._in(v, self.num_visible)
        ctx, st = self._ctx()
        B = v.shape[0]
        p = torch.empty(B, self.num_hidden, device=v.device, dtype=torch.float32)
        s = torch.empty_like(p) if sample else None
        rs = self._struct()
        ctx.check(ctx.lib.imdbn_up(ctx.handle, C.byref(rs), L.ptr(v), B, float(T), L.ptr(p), L.ptr(s),
                                   C.byref(rng) if rng is not None else None, draw, st), "imdbn_up")
        return p, s

    def _down(self, h, T=1.0, want_p=True, want_logits=False, sample=False, rng=None,
              draw_u=0, draw_cat=0):
        h = self._in(h, self.num_hidden)
        ctx, st = self._ctx()
        B = h.shape[0]
        mk = lambda: torch.empty(B, self.num_visible, device=h.device, dtype=torch.float32)
        p = mk() if want_p else None
        lg = mk() if want_logits else None
        s = mk() if sample else None
        rs = self._struct()
        ctx.check(ctx.lib.imdbn_down(ctx.handle, C.byref(rs), L.ptr(h), B, float(T), L.ptr(p), L.ptr(lg),
                                     L.ptr(s), C.byref(rng) if rng is not None else None,
                                     draw_u, draw_cat, st), "imdbn_down")
        return p, lg, s

    @torch.no_grad()
    def forward(self, v: torch.Tensor, T: float = 1.0) -> torch.Tensor:
        """p(h|v) = sigmoid((vW + b_h)/max(1e-6,T))  (rbm.py:81-92)."""
        return self._up(v, T)[0]

    @torch.no_grad()
    def _visible_logits(self, h: torch.Tensor, T: float = 1.0) -> torch.Tensor:
        """(hW^T + b_v)/max(1e-6,T)  (rbm.py:94-96)."""
        return self._down(h, T, want_p=False, want_logits=True)[1]

    @torch.no_grad()
    def visible_probs(self, h: torch.Tensor, T: float = 1.0) -> torch.Tensor:
        """p(v|h), softmax inside every softmax group  (rbm.py:98-116)."""
        return self._down(h, T)[0]

    @torch.no_grad()
    def sample_visible(self, v_prob: torch.Tensor) -> torch.Tensor:
        """Bernoulli sample, one-hot categorical inside softmax groups  (rbm.py:118-135)."""
        p = self._in(v_prob, self.num_visible)
        ctx, st = self._ctx()
        out = torch.empty_like(p)
        rs, rng = self._struct(), self._next_rng()
        ctx.check(ctx.lib.imdbn_sample_visible(ctx.handle, C.byref(rs), L.ptr(p), p.shape[0], L.ptr(out),
                                               C.byref(rng), 0, 1, st), "imdbn_sample_visible")
        return out

    @torch.no_grad()
    def backward(self, h: torch.Tensor, return_logits: bool = False) -> torch.Tensor:
        """Decoder pass  (rbm.py:137-151)."""
        return self._visible_logits(h) if return_logits else self.visible_probs(h)

    @torch.no_grad()
    def backward_sample(self, h: torch.Tensor) -> torch.Tensor:
        """rbm.py:153-156."""
        return self._down(h, 1.0, want_p=False, sample=True, rng=self._next_rng(), draw_u=0, draw_cat=1)[2]

    @torch.no_grad()
    def gibbs_step(self, v: torch.Tensor, sample_h: bool = True, sample_v: bool = True):
        """One v -> h -> v' step, returns (v_next, v_prob, h, h_prob)  (rbm.py:158-178).
        Draws: 0 = U[B,H], 1 = U[B,V], 2 = categorical."""
        rng = self._next_rng()
        h_prob, h_s = self._up(v, 1.0, sample=sample_h, rng=rng, draw=0)
        h = h_s if sample_h else h_prob
        v_prob, _, v_s = self._down(h, 1.0, sample=sample_v, rng=rng, draw_u=1, draw_cat=2)
        return (v_s if sample_v else v_prob), v_prob, h, h_prob

    # ------------------------------------------------------------------ CD-k
    def _update_struct(self, lr, mom, bsz, sparsity) -> L.UpdateStruct:
        return L.UpdateStruct(lr, mom, self.weight_decay, int(bool(sparsity)), self.sparsity_factor,
                              int(bsz))

    @torch.no_grad()
    def train_epoch(self, data: torch.Tensor, epoch: int, max_epochs: int, CD: int = 1):
        """One CD-k update on a minibatch; returns the reconstruction MSE as a 0-dim tensor
        (rbm.py:180-227).  With data parallelism enabled (``dist.enable``) ``data`` is this rank's
        shard and the statistics are summed over ranks before the (replicated) update."""
        data = self._in(data, self.num_visible)
        ctx, st = self._ctx()
        B = data.shape[0]
        lr, mom = self._hyper(epoch)
        dp = _dist.state()
        if dp is not None:
            self._dp_sync_params(dp)            # (collective, first call only) rank 0's parameters everywhere
        if dp is not None and dp.p2p:
            self._p2p_setup(dp)                 # (collective, first call only) re-homes W in peer-mapped memory
        rs = self._struct(training=True)
        loss = torch.empty((), device=data.device, dtype=torch.float32)
        self._n_updates = getattr(self, "_n_updates", 0) + 1
        if dp is None:
            rng = self._next_rng()
            upd = self._update_struct(lr, mom, B, self.sparsity)
            ctx.check(ctx.lib.imdbn_cd_train(ctx.handle, C.byref(rs), L.ptr(data), B, int(CD),
                                             C.byref(upd), C.byref(rng), L.ptr(loss), st),
                      "imdbn_cd_train")
            return loss
        self._train_dp(dp, ctx, st, rs, data, B, CD, lr, mom, None, loss)
        return loss

    def _train_dp(self, dp, ctx, st, rs, data, B, CD, lr, mom, pos_in, loss) -> None:
        """Sharded minibatch: local statistics -> cross-rank sum -> replicated update (dist.py)."""
        rng = self._next_rng(row0=dp.rank * B)
        stats = self._stats_buffer(ctx, rs)
        ctx.check(ctx.lib.imdbn_cd_stats(ctx.handle, C.byref(rs), L.ptr(data), B, int(CD),
                                         C.byref(rng), L.ptr(pos_in), L.ptr(stats), st), "imdbn_cd_stats")
        upd = self._update_struct(lr, mom, B * dp.world, self.sparsity)
        self._dp_apply(dp, ctx, rs, stats, upd, loss, st)

    # ------------------------------------------------------------------ data parallelism
    def _dp_sync_params(self, dp) -> None:
        """Replicated data parallelism needs identical parameters and momenta on every rank, and nothing else
        establishes that (the constructor draws W from each process's own generator): the first update under a
        given ``dist.enable()`` broadcasts rank 0's state.  Collective; once per (RBM, dist state)."""
        if self.__dict__.get("_dp_synced") is dp:
            return
        import torch.distributed as td
        self._sync_buffers()
        for t in (self.W.data, self.hid_bias.data, self.vis_bias.data, self.W_m, self.hb_m, self.vb_m):
            td.broadcast(t, src=td.get_global_rank(dp.group, 0) if dp.group is not None else 0, group=dp.group)
        self._dp_synced = dp

    def _p2p_setup(self, dp):
        """Peer-memory data parallelism (dist.py, csrc/dp_update.cuh): the statistics buffer and W live in
        symmetric memory that every rank of the box can load from / store to over NVLink."""
        st = self.__dict__.get("_p2p")
        if st is not None and st["dp"] is dp and st["w_ptr"] == self.W.data_ptr():
            return st
        V, H = self.num_visible, self.num_hidden
        if (V * H) % 4 or not self.W.is_cuda:
            self._p2p = None                                     # falls back to the NCCL all-reduce
            return None
        S, hS = dp.symm_empty(V * H + 2 * H + V + 1)
        Wflat, hW = dp.symm_empty(V * H)
        Wsym = Wflat.view(V, H)
        Wsym.copy_(self.W.data)
        self.W.data = Wsym
        peers = L.PeersStruct()
        peers.world, peers.rank = dp.world, dp.rank
        for r in range(dp.world):
            peers.stats[r] = hS.buffer_ptrs[r]
            peers.W[r] = hW.buffer_ptrs[r]
        if dp.multicast:                                         # NVSwitch multicast objects (NVLS)
            mcS, mcW = int(getattr(hS, "multicast_ptr", 0) or 0), int(getattr(hW, "multicast_ptr", 0) or 0)
            if mcS and mcW:
                peers.stats_mc, peers.W_mc = mcS, mcW
        st = dict(dp=dp, w_ptr=self.W.data_ptr(), S=S, hS=hS, hW=hW, peers=peers)
        self._p2p = st
        hW.barrier(channel=0)                                    # every copy is in place before anyone stores into it
        return st

    def _dp_apply(self, dp, ctx, rs, stats, upd, loss, st) -> None:
        """Make the local statistics global and apply the update on every rank (rbm.py:211-226)."""
        p2p = self.__dict__.get("_p2p") if dp.p2p else None
        if p2p is None:
            dp.all_reduce(stats)
            ctx.check(ctx.lib.imdbn_apply_update(ctx.handle, C.byref(rs), L.ptr(stats), C.byref(upd),
                                                 L.ptr(loss), st), "imdbn_apply_update")
            return
        p2p["hS"].barrier(channel=0)                             # all ranks' statistics are complete
        ctx.check(ctx.lib.imdbn_dp_update(ctx.handle, C.byref(rs), C.byref(p2p["peers"]), C.byref(upd),
                                          L.ptr(loss), st), "imdbn_dp_update")
        p2p["hS"].barrier(channel=1)                             # all slabs of W are in place everywhere

    def sync_momenta(self) -> None:
        """Peer-memory data parallelism keeps W_m only on the rank that owns each slab; this collective makes
        it whole on every rank (call on ALL ranks, e.g. before saving or before ``dist.disable()``)."""
        dp = _dist.state()
        if dp is None or self.__dict__.get("_p2p") is None:
            return
        flat = self.W_m.view(-1)
        q0, q1 = dp.slab(flat.numel() // 4)
        flat[:4 * q0].zero_()
        flat[4 * q1:].zero_()
        dp.all_reduce(flat)

    @torch.no_grad()
    def train_epoch_fwd(self, data: torch.Tensor, epoch: int, max_epochs: int, CD: int = 1,
                        next_data: Optional[torch.Tensor] = None, loss_out: Optional[torch.Tensor] = None):
        """``loss = train_epoch(data, ...); h = forward(data)`` as the iDBN training loop issues them
        (reference idbn.py:202-203), returned as ``(loss, h)``.  When ``next_data`` (the next minibatch of
        the same loader) is given, the forward pass also computes its positive hidden probabilities in the
        same pass over the updated ``W`` and keeps them for the next call, which then skips its first up
        pass: one 4*V*H-byte stream of ``W`` less per minibatch, bit-identical results.  ``loss_out`` = an
        fp32 scalar the loss is written to by the kernel itself: a CUDA tensor, or PINNED host memory (the
        device writes it over PCIe -- a device-to-host read-back without a copy operation in the stream)."""
        data = self._in(data, self.num_visible)
        dp = _dist.state()
        if dp is not None:
            return self._train_epoch_fwd_dp(dp, data, epoch, CD, next_data, loss_out)
        ctx, st = self._ctx()
        B = data.shape[0]
        lr, mom = self._hyper(epoch)
        rs = self._struct(training=True)
        if loss_out is None:
            loss = torch.empty((), device=data.device, dtype=torch.float32)
        else:
            if loss_out.dtype != torch.float32 or loss_out.numel() != 1 or not (
                    loss_out.is_cuda or loss_out.is_pinned()):
                raise ValueError("loss_out must be one fp32 element on the device or in pinned host memory")
            loss = loss_out
        rng = self._next_rng()
        upd = self._update_struct(lr, mom, B, self.sparsity)
        cached = self.__dict__.pop("_pos_cache", None)
        pos_in = None
        if cached is not None:
            key, pos = cached
            if key == self._pos_key(data):
                pos_in = pos
        nxt, Bn = None, 0
        if next_data is not None:
            nxt = self._in(next_data, self.num_visible)
            Bn = nxt.shape[0]
        fwd = torch.empty(B + Bn, self.num_hidden, device=data.device, dtype=torch.float32)
        ctx.check(ctx.lib.imdbn_cd_train_fwd(ctx.handle, C.byref(rs), L.ptr(data), B, int(CD), C.byref(upd),
                                             C.byref(rng), L.ptr(loss), L.ptr(pos_in), L.ptr(nxt), Bn,
                                             L.ptr(fwd), st), "imdbn_cd_train_fwd")
        self._n_updates = getattr(self, "_n_updates", 0) + 1
        if nxt is not None:
            self._pos_cache = (self._pos_key(nxt), fwd[B:])
        return loss, fwd[:B]

    def _train_epoch_fwd_dp(self, dp, data, epoch, CD, next_data, loss_out):
        """``train_epoch_fwd`` on a sharded minibatch: the cached positive phase and the fused
        forward-plus-next-positive pass work per shard exactly as on one GPU; only the update is global."""
        ctx, st = self._ctx()
        B = data.shape[0]
        lr, mom = self._hyper(epoch)
        self._dp_sync_params(dp)
        if dp.p2p:
            self._p2p_setup(dp)
        rs = self._struct(training=True)
        loss = torch.empty((), device=data.device, dtype=torch.float32)
        cached = self.__dict__.pop("_pos_cache", None)
        pos_in = None
        if cached is not None and cached[0] == self._pos_key(data):
            pos_in = cached[1]
        self._train_dp(dp, ctx, st, rs, data, B, CD, lr, mom, pos_in, loss)
        self._n_updates = getattr(self, "_n_updates", 0) + 1
        if loss_out is not None:
            loss_out.copy_(loss, non_blocking=True)
        if next_data is None:
            return loss, self.forward(data)
        nxt = self._in(next_data, self.num_visible)
        fwd = self.forward(torch.cat([data, nxt], 0))            # one pass over the updated W for both
        self._pos_cache = (self._pos_key(nxt), fwd[B:])
        return loss, fwd[:B]

    def _pos_key(self, x: torch.Tensor):
        """Identity of (input buffer, parameter state) under which cached positive probabilities hold."""
        return (x.data_ptr(), tuple(x.shape), x._version, self.W.data_ptr(), self.W._version,
                self.hid_bias._version, getattr(self, "_n_updates", 0))

    def _stats_buffer(self, ctx, rs) -> torch.Tensor:
        p2p = self.__dict__.get("_p2p")
        if p2p is not None and _dist.state() is not None and _dist.state().p2p:
            return p2p["S"]                                      # peer-mapped statistics buffer
        n = int(ctx.lib.imdbn_stats_size(C.byref(rs)))
        buf = getattr(self, "_stats_buf", None)
        if buf is None or buf.numel() != n or buf.device != self.W.device:
            buf = torch.empty(n, device=self.W.device, dtype=torch.float32)
            self._stats_buf = buf
        return buf

    # ------------------------------------------------------------------ schedules
    def _lin_schedule(self, t, t_max, start, end):
        """rbm.py:229-234."""
        return _schedule(t, t_max, start, end)

    def _hot_steps(self, n_steps, hot_frac):
        """rbm.py:236-238 (the reference computes this and never uses it)."""
        return int(max(0, min(n_steps, round(hot_frac * n_steps))))

    # ------------------------------------------------------------------ conditional inference
    def _run_chain(self, kind, v_known, known_mask, n_steps, tables=None, mu=None, sample_h=False,
                   sample_v=False, final_free=False, v_init=None, draw0=0, rng=None,
                   want_vprob=False, clamp_prefix=-1, clamp_suffix=-1):
        vk = self._in(v_known, self.num_visible)
        km = self._in(known_mask, self.num_visible)
        if km.shape != vk.shape:
            raise RuntimeError("known_mask must have the shape of v_known")
        ctx, st = self._ctx()
        B = vk.shape[0]
        ch = L.ChainStruct()
        ch.kind, ch.n_steps = kind, int(n_steps)
        ch.v_known, ch.known_mask = vk.data_ptr(), km.data_ptr()
        vi = None
        if v_init is not None:
            vi = self._in(v_init, self.num_visible)
            ch.v_init = vi.data_ptr()
        keep = []
        if tables is not None:
            for name, arr in zip(("T", "sigma", "eta"), tables):
                buf = (C.c_float * max(1, len(arr)))(*arr)
                keep.append(buf)
                setattr(ch, name, C.cast(buf, L.c_float_p))
        mu_t = None
        if mu is not None:
            mu_t = L.f32c(mu, vk.device)
            if mu_t.dim() != 2 or mu_t.shape[0] != B or mu_t.shape[1] > self.num_visible:
                raise RuntimeError("mu_k must be [B, Dz] with Dz <= num_visible")
            ch.mu, ch.Dz = mu_t.data_ptr(), mu_t.shape[1]
        ch.sample_h, ch.sample_v, ch.final_free_sweep = int(sample_h), int(sample_v), int(final_free)
        ch.draw0 = draw0 & 0xFFFFFFFF
        ch.clamp_prefix = int(clamp_prefix)
        ch.clamp_suffix = int(clamp_suffix)
        out = torch.empty_like(vk)
        vprob = torch.empty_like(vk) if want_vprob else None
        rs = self._struct()
        rng = rng if rng is not None else self._next_rng()
        ctx.check(ctx.lib.imdbn_run_chain(ctx.handle, C.byref(rs), C.byref(ch), B, L.ptr(out), L.ptr(vprob),
                                          C.byref(rng), st), "imdbn_run_chain")
        return (out, vprob) if want_vprob else out

    @torch.no_grad()
    def noisy_meanfield_annealed(self, v_known: torch.Tensor, known_mask: torch.Tensor,
                                 n_steps: int = 72, T0: float = 3.0, T1: float = 1.0,
                                 sigma0: float = 0.9, hot_frac: float = 0.7, sharpen_last: int = 3,
                                 T_cold_plus: float = 0.9, *, clamp_suffix: int = -1):
        """Noisy mean-field annealing with optional mu-pull (``self._mu_pull``), rbm.py:300-367.
        The whole chain is one persistent kernel.  ``hot_frac`` is accepted and, as in the
        reference, has no effect.  Draws: 0 = U[B,V] init; step t: 1+2t = N[B,H], 2+2t = N[B,V].
        ``clamp_suffix = Dz`` (keyword-only extension) is the caller's promise that ``known_mask`` is 1 on
        exactly the columns from Dz on (TXT->IMG inference): the large-batch chain then does no work on them."""
        n = int(n_steps)
        Ts, Ss, Es = [], [], []
        pull = getattr(self, "_mu_pull", None)
        eta0 = float(pull.get("eta0", 0.15)) if pull is not None else 0.0
        for t in range(n):
            Tt = _schedule(t, n, T0, T1)
            if (n - t) <= max(1, int(sharpen_last)):
                Tt = T_cold_plus
            frac = max(0.0, 1.0 - (t / max(1, n - 1)))
            Ts.append(max(1e-6, Tt)); Ss.append(sigma0 * frac); Es.append(eta0 * frac)
        return self._run_chain(L.CHAIN_NOISY_MF, v_known, known_mask, n, tables=(Ts, Ss, Es),
                               mu=pull["mu_k"] if pull is not None else None, clamp_suffix=clamp_suffix)

    @torch.no_grad()
    def conditional_gibbs(self, v_known: torch.Tensor, known_mask: torch.Tensor, n_steps: int = 30,
                          sample_h: bool = False, sample_v: bool = False, *,
                          clamp_prefix: int = -1) -> torch.Tensor:
        """n clamped sweeps then one un-clamped sweep, rbm.py:369-400.  Draws: 0 = U[B,V] init;
        step t: 1+3t = U[B,H], 2+3t = U[B,V], 3+3t = categorical.
        ``clamp_prefix = Dz`` (keyword-only extension) is the caller's promise that ``known_mask`` is 1 on
        exactly the first Dz columns: with a single trailing softmax group the chain then runs in the
        label-only kernel (IMG->TXT inference of ``iMDBN._cross_reconstruct``)."""
        return self._run_chain(L.CHAIN_COND_GIBBS, v_known, known_mask, int(n_steps),
                               sample_h=sample_h, sample_v=sample_v, final_free=True,
                               clamp_prefix=clamp_prefix)

    @torch.no_grad()
    def conditional_gibbs_annealed(self, v_known: torch.Tensor, known_mask: torch.Tensor,
                                   n_steps: int = 40, T0: float = 2.5, T1: float = 1.0,
                                   sample_h_until: int = 20, sample_v_every: int = 0,
                                   final_meanfield: bool = True):
        """Annealed conditional Gibbs, rbm.py:240-298 (not called anywhere in the reference; kept for
        API parity, sequenced on the host over the pass kernels).  Draws: 0 = U[B,V] init; step t:
        1+3t = U[B,H], 2+3t = U[B,V], 3+3t = categorical."""
        vk = self._in(v_known, self.num_visible)
        km = self._in(known_mask, self.num_visible)
        rng = self._next_rng()
        u0 = random_field(rng, 0, vk.shape[0], vk.shape[1], vk.device)
        v = vk * km + (1 - km) * u0
        hot = int(max(0, min(n_steps, sample_h_until)))
        for t in range(int(n_steps)):
            Tt = _schedule(t, n_steps, T0, T1)
            if (n_steps - t) <= 3:
                Tt = min(0.9, Tt)
            h_prob, h_s = self._up(v, Tt, sample=t < hot, rng=rng, draw=1 + 3 * t)
            h = h_s if t < hot else h_prob
            samp_v = (t < hot) and (sample_v_every > 0) and (t % sample_v_every == 0)
            p, _, s = self._down(h, Tt, sample=samp_v, rng=rng, draw_u=2 + 3 * t, draw_cat=3 + 3 * t)
            v = (s if samp_v else p) * (1 - km) + vk * km
        if final_meanfield:
            v = self.visible_probs(self.forward(v, T=1.0), T=1.0) * (1 - km) + vk * km
        return v

    @torch.no_grad()
    def train_epoch_clamped(self, v_known: torch.Tensor, known_mask: torch.Tensor, epoch: int,
                            max_epochs: int, CD: int = 1, cond_init_steps: int = 50,
                            sample_h: bool = True, sample_v: bool = False,
                            reclamp_negative: bool = True, aux_lr_mult: float = 0.3,
                            use_noisy_init: bool = True):
        """Auxiliary clamped CD, rbm.py:402-483; returns mean((v+ - v-)^2) as a 0-dim tensor."""
        vk = self._in(v_known, self.num_visible)
        km = self._in(known_mask, self.num_visible)
        ctx, st = self._ctx()
        B = vk.shape[0]
        lr, mom = self._hyper(epoch, aux_lr_mult)
        dp = _dist.state()
        if dp is not None:
            self._dp_sync_params(dp)
        if dp is not None and dp.p2p:
            self._p2p_setup(dp)
        rs = self._struct(training=True)
        cfg = L.ClampedCfgStruct(int(CD), int(cond_init_steps), int(sample_h), int(sample_v),
                                 int(reclamp_negative), int(use_noisy_init))
        loss = torch.empty((), device=vk.device, dtype=torch.float32)
        self._n_updates = getattr(self, "_n_updates", 0) + 1
        if dp is None:
            rng = self._next_rng()
            upd = self._update_struct(lr, mom, B, False)
            ctx.check(ctx.lib.imdbn_cd_train_clamped(ctx.handle, C.byref(rs), L.ptr(vk), L.ptr(km), B,
                                                     C.byref(cfg), C.byref(upd), C.byref(rng),
                                                     L.ptr(loss), st), "imdbn_cd_train_clamped")
            return loss
        rng = self._next_rng(row0=dp.rank * B)
        stats = self._stats_buffer(ctx, rs)
        ctx.check(ctx.lib.imdbn_cd_clamped_stats(ctx.handle, C.byref(rs), L.ptr(vk), L.ptr(km), B,
                                                 C.byref(cfg), C.byref(rng), L.ptr(stats), st),
                  "imdbn_cd_clamped_stats")
        upd = self._update_struct(lr, mom, B * dp.world, False)
        self._dp_apply(dp, ctx, rs, stats, upd, loss, st)
        return loss

    # ------------------------------------------------------------------ pickling
    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("_stats_buf", None)       # scratch, not model state
        state.pop("_pos_cache", None)
        state.pop("_p2p", None)             # peer-memory handles
        state.pop("_dp_synced", None)
        return state


def rbm_free_energy(rbm: RBM, v: torch.Tensor) -> torch.Tensor:
    """F(v) = -v.b_v - sum_j softplus(b_h + vW)_j  (reference utils/energy_utils.py:18-28).
    Install as ``RBM.free_energy = rbm_free_energy`` to make best-of-K in
    ``iMDBN._cross_reconstruct`` live (the reference ships without it)."""
    v = rbm._in(v, rbm.num_visible)
    ctx, st = rbm._ctx()
    out = torch.empty(v.shape[0], device=v.device, dtype=torch.float32)
    rs = rbm._struct()
    ctx.check(ctx.lib.imdbn_free_energy(ctx.handle, C.byref(rs), L.ptr(v), v.shape[0], L.ptr(out), st),
              "imdbn_free_energy")
    return out


def random_field(rng: L.RngStruct, draw: int, rows: int, cols: int, device, normal: bool = False):
    """Materialise one tensor of the random field (tests, host-sequenced chains)."""
    out = torch.empty(rows, cols, device=device, dtype=torch.float32)
    ctx, st = L.context_for(out)
    ctx.check(ctx.lib.imdbn_random_field(ctx.handle, C.byref(rng), draw, int(normal), rows, cols,
                                         L.ptr(out), st), "imdbn_random_field")
    return out
