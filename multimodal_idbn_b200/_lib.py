"""ctypes binding of ``libimdbn_b200.so`` (C ABI declared in ``include/imdbn_b200.h``).

There is no CPU fallback: if the shared library is missing or a tensor is not on a CUDA device
the call raises.  Build the library with ``python -m multimodal_idbn_b200.build`` (or
``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Dict, Optional, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libimdbn_b200.so")

MAX_GROUPS = 4
PREC_FP32, PREC_TF32, PREC_TF32X2 = 0, 1, 2
CHAIN_NOISY_MF, CHAIN_COND_GIBBS = 0, 1
KERNEL_UP, KERNEL_DOWN, KERNEL_STATS, KERNEL_CHAIN, KERNEL_PACK = 0, 1, 2, 3, 4

c_float_p = C.POINTER(C.c_float)


class RbmStruct(C.Structure):
    _fields_ = [("W", C.c_void_p), ("hb", C.c_void_p), ("vb", C.c_void_p),
                ("Wm", C.c_void_p), ("hbm", C.c_void_p), ("vbm", C.c_void_p),
                ("V", C.c_int32), ("H", C.c_int32), ("ngroups", C.c_int32),
                ("group_start", C.c_int32 * MAX_GROUPS), ("group_end", C.c_int32 * MAX_GROUPS)]


class RngStruct(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("stream", C.c_uint32), ("row0", C.c_uint32)]


class UpdateStruct(C.Structure):
    _fields_ = [("lr", C.c_float), ("momentum", C.c_float), ("weight_decay", C.c_float),
                ("sparsity", C.c_int32), ("sparsity_target", C.c_float),
                ("batch_global", C.c_int32)]


MAX_PEERS = 8


class PeersStruct(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32),
                ("stats", C.c_void_p * MAX_PEERS), ("W", C.c_void_p * MAX_PEERS),
                ("stats_mc", C.c_void_p), ("W_mc", C.c_void_p)]


class ChainStruct(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_steps", C.c_int32),
                ("v_known", C.c_void_p), ("known_mask", C.c_void_p), ("v_init", C.c_void_p),
                ("T", c_float_p), ("sigma", c_float_p), ("eta", c_float_p),
                ("mu", C.c_void_p), ("Dz", C.c_int32),
                ("sample_h", C.c_int32), ("sample_v", C.c_int32),
                ("final_free_sweep", C.c_int32), ("draw0", C.c_uint32), ("clamp_prefix", C.c_int32),
                ("clamp_suffix", C.c_int32)]


class ClampedCfgStruct(C.Structure):
    _fields_ = [("k", C.c_int32), ("cond_init_steps", C.c_int32), ("sample_h", C.c_int32),
                ("sample_v", C.c_int32), ("reclamp_negative", C.c_int32),
                ("use_noisy_init", C.c_int32)]


# name -> (restype, argtypes); the single source for the loader AND the symbol test
_P = C.c_void_p
_I = C.c_int
_U = C.c_uint32
_F = C.c_float
SIGNATURES = {
    "imdbn_abi_version": (_I, []),
    "imdbn_ctx_create": (_I, [C.POINTER(_P), _I]),
    "imdbn_ctx_destroy": (None, [_P]),
    "imdbn_last_error": (C.c_char_p, [_P]),
    "imdbn_set_precision": (_I, [_P, _I]),
    "imdbn_launch_count": (C.c_int64, [_P]),
    "imdbn_copy_async": (_I, [_P, _P, C.c_size_t, _P]),
    "imdbn_profile_enable": (_I, [_P, _I]),
    "imdbn_profile_read": (_I, [_P, _I, _I, _I, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "imdbn_up": (_I, [_P, C.POINTER(RbmStruct), _P, _I, _F, _P, _P, C.POINTER(RngStruct), _U, _P]),
    "imdbn_down": (_I, [_P, C.POINTER(RbmStruct), _P, _I, _F, _P, _P, _P, C.POINTER(RngStruct),
                        _U, _U, _P]),
    "imdbn_sample_visible": (_I, [_P, C.POINTER(RbmStruct), _P, _I, _P, C.POINTER(RngStruct), _U, _U, _P]),
    "imdbn_free_energy": (_I, [_P, C.POINTER(RbmStruct), _P, _I, _P, _P]),
    "imdbn_cd_train": (_I, [_P, C.POINTER(RbmStruct), _P, _I, _I, C.POINTER(UpdateStruct),
                            C.POINTER(RngStruct), _P, _P]),
    "imdbn_cd_train_fwd": (_I, [_P, C.POINTER(RbmStruct), _P, _I, _I, C.POINTER(UpdateStruct),
                                C.POINTER(RngStruct), _P, _P, _P, _I, _P, _P]),
    "imdbn_cd_stats": (_I, [_P, C.POINTER(RbmStruct), _P, _I, _I, C.POINTER(RngStruct), _P, _P, _P]),
    "imdbn_stats_size": (C.c_int64, [C.POINTER(RbmStruct)]),
    "imdbn_apply_update": (_I, [_P, C.POINTER(RbmStruct), _P, C.POINTER(UpdateStruct), _P, _P]),
    "imdbn_set_sm_limit": (_I, [_P, _I]),
    "imdbn_idbn_train_step": (_I, [_P, _P, _I, C.POINTER(RbmStruct), C.POINTER(UpdateStruct), C.POINTER(RngStruct), _P, _I,
                                   _I, _P, _P, _I, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _I, _P, _P, _P, _I]),
    "imdbn_sm_partition": (_I, [_I, _I, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "imdbn_dp_update": (_I, [_P, C.POINTER(RbmStruct), C.POINTER(PeersStruct), C.POINTER(UpdateStruct), _P, _P]),
    "imdbn_class_free_energies": (_I, [_P, C.POINTER(RbmStruct), _P, _I, _I, _P, _P]),
    "imdbn_trace_img2txt": (_I, [_P, C.POINTER(RbmStruct), _P, _I, _I, _P, _I, _P, _P]),
    "imdbn_assoc_stats": (_I, [_P, C.POINTER(RbmStruct), _P, _P, _P, _P, _I, _P, _P]),
    "imdbn_run_chain": (_I, [_P, C.POINTER(RbmStruct), C.POINTER(ChainStruct), _I, _P, _P,
                             C.POINTER(RngStruct), _P]),
    "imdbn_cd_train_clamped": (_I, [_P, C.POINTER(RbmStruct), _P, _P, _I, C.POINTER(ClampedCfgStruct),
                                    C.POINTER(UpdateStruct), C.POINTER(RngStruct), _P, _P]),
    "imdbn_cd_clamped_stats": (_I, [_P, C.POINTER(RbmStruct), _P, _P, _I, C.POINTER(ClampedCfgStruct),
                                    C.POINTER(RngStruct), _P, _P]),
    "imdbn_best_of_k": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P]),
    "imdbn_class_stats": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P]),
    "imdbn_random_field": (_I, [_P, C.POINTER(RngStruct), _U, _I, _I, _I, _P, _P]),
}

_lib = None
_lib_lock = threading.Lock()


def load_library():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension is not built and there is no CPU "
                "fallback.  Run `python -m multimodal_idbn_b200.build`.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.imdbn_abi_version() != 1:
            raise RuntimeError("libimdbn_b200.so ABI version mismatch; rebuild")
        _lib = lib
    return _lib


class Context:
    """One library context (workspace arena) per (device, stream)."""

    def __init__(self, device_index: int):
        self.lib = load_library()
        h = C.c_void_p()
        with torch.cuda.device(device_index):
            rc = self.lib.imdbn_ctx_create(C.byref(h), int(device_index))
        if rc != 0:
            raise RuntimeError(f"imdbn_ctx_create(device={device_index}) failed with code {rc}")
        self.handle = h
        self.device_index = device_index
        self.precision = PREC_FP32

    def check(self, rc: int, what: str):
        if rc == 0:
            return
        msg = self.lib.imdbn_last_error(self.handle)
        msg = msg.decode() if msg else ""
        if rc < 0:
            raise ValueError(f"{what}: {msg} (code {rc})")
        raise RuntimeError(f"{what}: CUDA error {rc}: {msg}")

    def set_precision(self, prec: int):
        if prec != self.precision:
            self.check(self.lib.imdbn_set_precision(self.handle, int(prec)), "imdbn_set_precision")
            self.precision = prec

    def set_sm_limit(self, n_sms: int):
        """At most ``n_sms`` SMs for the persistent tensor-core kernels of this context (0 = all)."""
        self.check(self.lib.imdbn_set_sm_limit(self.handle, int(n_sms)), "imdbn_set_sm_limit")

    def profile(self, enable: bool):
        self.check(self.lib.imdbn_profile_enable(self.handle, int(enable)), "imdbn_profile_enable")

    def profile_read(self, kind: int, V: int, H: int):
        ms, n = C.c_double(), C.c_int64()
        self.check(self.lib.imdbn_profile_read(self.handle, kind, V, H, C.byref(ms), C.byref(n)),
                   "imdbn_profile_read")
        return ms.value, n.value

    def launch_count(self) -> int:
        return int(self.lib.imdbn_launch_count(self.handle))

    def __del__(self):
        try:
            if self.handle:
                self.lib.imdbn_ctx_destroy(self.handle)
        except Exception:
            pass


_contexts: Dict[Tuple[int, int], Context] = {}
_precision = PREC_FP32


def current_precision() -> int:
    return _precision


def set_precision(name: str):
    """'fp32' (parity mode, FFMA), 'tf32' (tcgen05 tensor-core passes, 1e-3) or 'tf32x2' (exact mode on the tensor
    cores: hi + lo tf32 terms of every operand, same parity bars as 'fp32')."""
    global _precision
    table = {"fp32": PREC_FP32, "tf32": PREC_TF32, "tf32x2": PREC_TF32X2}
    if name not in table:
        raise ValueError(f"precision must be one of {sorted(table)}")
    _precision = table[name]


def get_precision() -> str:
    return {PREC_FP32: "fp32", PREC_TF32: "tf32", PREC_TF32X2: "tf32x2"}[_precision]


def context_for(t: torch.Tensor) -> Tuple[Context, int]:
    """Context + raw stream handle for the device of ``t`` and torch's current stream."""
    if not t.is_cuda:
        raise RuntimeError(
            "multimodal_idbn_b200 runs on CUDA only (no CPU fallback); got a tensor on "
            f"'{t.device}'. Move the model and data to a B200 device.")
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    stream = torch.cuda.current_stream(idx).cuda_stream
    key = (idx, int(stream))
    ctx = _contexts.get(key)
    if ctx is None:
        ctx = Context(idx)
        _contexts[key] = ctx
    ctx.set_precision(_precision)
    return ctx, int(stream)


_partitions = {}


def sm_partition(device_index: int, small_sms: int):
    """``(stream_big, stream_small, n_big, n_small)`` -- raw handles of two streams on disjoint SM sets (CUDA green
    contexts, ``imdbn_sm_partition``), or None when the driver cannot provide them.  One partition per (device, size)."""
    # Nsight Compute cannot profile kernels of green contexts ("Failed to prepare kernel for profiling"): profiling
    # runs use IMDBN_NO_PARTITION=1 (layer pipelining then falls back to an ordinary side stream)
    if os.environ.get("IMDBN_NO_PARTITION") or os.environ.get("CUDA_INJECTION64_PATH"):
        return None
    key = (device_index, int(small_sms))
    if key not in _partitions:
        lib = load_library()
        big, small = C.c_void_p(), C.c_void_p()
        nb, ns = C.c_int(), C.c_int()
        rc = lib.imdbn_sm_partition(int(device_index), int(small_sms), C.byref(big), C.byref(small), C.byref(nb),
                                    C.byref(ns))
        _partitions[key] = (big.value, small.value, nb.value, ns.value) if rc == 0 and big.value else None
    return _partitions[key]


def context_for_stream(device_index: int, stream_handle: int) -> "Context":
    """Context bound to a raw stream handle (streams that torch did not create)."""
    key = (device_index, int(stream_handle))
    ctx = _contexts.get(key)
    if ctx is None:
        ctx = Context(device_index)
        _contexts[key] = ctx
    ctx.set_precision(_precision)
    return ctx


_private_contexts = []


def private_context(device_index: int) -> "Context":
    """A context of its own (workspace, SM limit) that no other caller on the same stream shares; counted by
    ``total_launches``."""
    import weakref
    ctx = Context(device_index)
    ctx.set_precision(_precision)
    _private_contexts.append(weakref.ref(ctx))
    return ctx


def total_launches() -> int:
    n = sum(c.launch_count() for c in _contexts.values())
    for ref in _private_contexts:
        c = ref()
        if c is not None:
            n += c.launch_count()
    return n


def ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def f32c(t: torch.Tensor, device) -> torch.Tensor:
    """fp32, contiguous, on ``device`` (no copy when already so); never tracks gradients."""
    t = t.detach()
    if t.device != device or t.dtype != torch.float32:
        t = t.to(device=device, dtype=torch.float32, non_blocking=True)
    return t.contiguous()
