"""Host (numpy) statement of the counter-based random field shared by the oracle and the CUDA kernels.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product package
(``multimodal_idbn_b200``); only ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of
``bench.py`` use it.

The reference (``imdbn/models/rbm.py``) draws whole tensors with ``torch.rand_like`` /
``torch.randn_like`` / ``Categorical.sample`` in a fixed order (rbm.py:125,131,203,208,333,346,352,
392,395,462).  Parity needs both sides to consume *identical* numbers, so the random numbers are
defined as a pure function of an address instead of a stream position:

    x[0..3] = Philox4x32-10( key = (seed_lo, seed_hi),
                             counter = (col, global_row, draw, stream) )

* ``stream``  : one number per API call (``train_epoch`` #17 of this RBM, ...).
* ``draw``    : which tensor inside that call (documented next to each oracle function).
* ``global_row`` / ``col`` : element address, so a batch sharded over ranks sees the same numbers.

uniform  u  = (x0 >> 8) * 2^-24                      in [0, 1)   (what ``rand_like`` would give)
normal   n  = sqrt(-2 ln u1) * cos(2 pi u2)          with u1 = ((x0 >> 8) + 1) * 2^-24 in (0, 1],
                                                     u2 = (x1 >> 8) * 2^-24
categorical: one uniform per (row, group): col = group index.

The CUDA side implements the same function in ``csrc/philox.cuh``.
"""
from __future__ import annotations

import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)
_INV24 = np.float32(1.0 / 16777216.0)
_TWO_PI = np.float32(6.283185307179586)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al., SC'11).  All inputs broadcastable uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(
        np.asarray(c0, dtype=np.uint32), np.asarray(c1, dtype=np.uint32),
        np.asarray(c2, dtype=np.uint32), np.asarray(c3, dtype=np.uint32))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for r in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & _MASK).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & _MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            if r != 9:
                k0 = np.uint32(k0 + _W0)
                k1 = np.uint32(k1 + _W1)
    return c0, c1, c2, c3


class RandomField:
    """The random numbers of ONE API call (fixed ``seed`` and ``stream``)."""

    def __init__(self, seed: int, stream: int):
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.stream = int(stream) & 0xFFFFFFFF
        self.k0 = self.seed & 0xFFFFFFFF
        self.k1 = (self.seed >> 32) & 0xFFFFFFFF

    def _raw(self, draw: int, rows: int, c_lo: int, c_hi: int, row0: int = 0):
        """Philox outputs for counters (c, row, draw, stream), c in [c_lo, c_hi] -> uint32 [rows, c_hi-c_lo+1, 4]."""
        r = (np.arange(rows, dtype=np.uint32) + np.uint32(row0))[:, None]
        c = np.arange(c_lo, c_hi + 1, dtype=np.uint32)[None, :]
        return np.stack(philox4x32_10(c, r, np.uint32(draw), np.uint32(self.stream), self.k0, self.k1), axis=-1)

    def uniform(self, draw: int, rows: int, cols: int, row0: int = 0, col0: int = 0) -> np.ndarray:
        """One Philox call serves FOUR consecutive columns: column j reads word j & 3 of counter j >> 2."""
        j = np.arange(col0, col0 + cols)
        x = self._raw(draw, rows, int(j[0]) >> 2, int(j[-1]) >> 2, row0)
        w = x[:, (j >> 2) - (int(j[0]) >> 2), j & 3]
        return (w >> np.uint32(8)).astype(np.float32) * _INV24

    def normal(self, draw: int, rows: int, cols: int, row0: int = 0, col0: int = 0) -> np.ndarray:
        """One Philox call serves TWO consecutive columns: column j reads words 2(j & 1), 2(j & 1) + 1 of counter
        j >> 1 (Box-Muller, cosine branch)."""
        j = np.arange(col0, col0 + cols)
        x = self._raw(draw, rows, int(j[0]) >> 1, int(j[-1]) >> 1, row0)
        idx = (j >> 1) - (int(j[0]) >> 1)
        a, b = x[:, idx, 2 * (j & 1)], x[:, idx, 2 * (j & 1) + 1]
        u1 = ((a >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * _INV24
        u2 = (b >> np.uint32(8)).astype(np.float32) * _INV24
        rad = np.sqrt(np.float32(-2.0) * np.log(u1, dtype=np.float32), dtype=np.float32)
        return (rad * np.cos(_TWO_PI * u2, dtype=np.float32)).astype(np.float32)

    def cat_uniform(self, draw: int, rows: int, group: int, row0: int = 0) -> np.ndarray:
        """One uniform per row for softmax group number ``group`` -> shape [rows]."""
        return self.uniform(draw, rows, 1, row0=row0, col0=group)[:, 0]


# Known-answer vectors of Philox4x32-10 from the Random123 distribution (kat_vectors):
#   counter, key -> output
KAT = [
    ((0x00000000,) * 4, (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]
