"""CPU oracle for the training-side ray-batch sampler.  TEST INFRASTRUCTURE ONLY (only ``tests/`` may import it).

Restates, in numpy integer arithmetic, the keyed permutation ``csrc/voxe_sampler.cu`` evaluates on the device (integer work:
the comparison is bit-exact), and what a selection means in the reference's terms: index i of
``collate_rays([flatten_rays(cast_rays(intrinsics, pose_b)) for b ...])`` / ``images.permute(0,2,3,1).reshape(-1, C)``
(thre3d_atom/modules/trainers.py:290-308; thre3d_atom/rendering/volumetric/utils/misc.py:12-50, 126-138).  The reference
itself draws ``torch.randperm(N)[:k]``; its contract -- k distinct uniformly distributed rows -- is what the tests check.

PARITY UNPINNED for the drawn numbers: no golden vector of the reference can pin them (they are ``torch.randperm``'s), so
this file is pinned only in the other direction -- the kernel must equal it bit for bit, and it must be a permutation.  The
meaning of an index (which ray, which pixel) is pinned against the reference's ``cast_rays`` through
``tests/golden/cameras.npz`` (tests/test_abi_and_api.py) and tests/test_sampler.py.
"""
from __future__ import annotations

import numpy as np

M32 = 0xFFFFFFFF


def pcg(v: int) -> int:
    state = (v * 747796405 + 2891336453) & M32
    word = (((state >> ((state >> 28) + 4)) ^ state) * 277803737) & M32
    return ((word >> 22) ^ word) & M32


class Permutation:
    """Cycle-walking 6-round balanced Feistel network over [0, 4^half_bits) >= [0, n), PCG hash as round function."""

    def __init__(self, n: int, seed: int, offset: int):
        self.n = n
        self.half_bits = 1
        while self.half_bits < 32 and (1 << (2 * self.half_bits)) < n:
            self.half_bits += 1
        k = pcg((seed & M32) ^ pcg((seed >> 32) & M32))
        k = pcg(k ^ (offset & M32))
        k = pcg(k ^ ((offset >> 32) & M32))
        self.key = []
        for j in range(6):
            k = pcg((k + 0x9E3779B9 * (j + 1)) & M32)
            self.key.append(k)

    def _network(self, v: int) -> int:
        mask = M32 if self.half_bits >= 32 else (1 << self.half_bits) - 1
        left, right = (v >> self.half_bits) & mask, v & mask
        for k in self.key:
            left, right = right, left ^ (pcg(right ^ k) & mask)
        return (left << self.half_bits) | right

    def at(self, i: int) -> int:
        v = self._network(i)
        while v >= self.n:
            v = self._network(v)
        return v

    def head(self, k: int) -> np.ndarray:
        return np.array([self.at(i) for i in range(k)], dtype=np.int64)
