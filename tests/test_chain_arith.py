"""The cursor replay's quotient trick (csrc/kernels_chain.cuh, chain_step_tdfa32), restated in numpy: the number of
times a record is returned is floor(gap / L) + 1; the kernel estimates the quotient as mulhi(gap, M) with
M = trunc(2^32 / L) taken from an APPROXIMATE float reciprocal and repairs it with one +1 / -1 step.  That is exact
as long as the estimate is within 1 of the true quotient -- checked here for gaps below 2^22 (the kernel's domain)
with the multiplier perturbed by up to 2 ulp in both directions (the stated error of __fdividef)."""
import numpy as np


def _corrected_quotient(gap, L, M):
    qf = (gap.astype(np.uint64) * M.astype(np.uint64)) >> np.uint64(32)
    rem = gap.astype(np.int64) - qf.astype(np.int64) * L.astype(np.int64)
    qf = qf.astype(np.int64) + (rem >= L.astype(np.int64)) - (rem < 0)
    return qf


def _multiplier(L, ulps):
    q = np.float32(4294967296.0) / L.astype(np.float32)
    for _ in range(abs(ulps)):
        q = np.nextafter(q, np.float32(np.inf if ulps > 0 else 0), dtype=np.float32)
    return np.minimum(np.floor(q.astype(np.float64)), 4294967295.0).astype(np.uint64)   # __float2uint_rz saturates


def test_quotient_estimate_is_repaired_exactly():
    rng = np.random.default_rng(11)
    n = 400_000
    L = np.concatenate([rng.integers(1, 1 << 22, size=n), rng.integers(1, 300, size=n), np.array([1, 2, 3, 7, (1 << 22) - 1])]).astype(np.uint32)
    gap = rng.integers(0, 1 << 22, size=L.size).astype(np.uint32)
    # adversarial gaps: multiples of L and their neighbours
    k = rng.integers(0, 1 << 12, size=L.size).astype(np.uint64)
    near = np.minimum(k * L.astype(np.uint64) + rng.integers(0, 3, size=L.size).astype(np.uint64), (1 << 22) - 1).astype(np.uint32)
    for g in (gap, near, np.minimum(near + L - 1, (1 << 22) - 1).astype(np.uint32)):
        want = g.astype(np.int64) // L.astype(np.int64)
        for ulps in (-2, -1, 0, 1, 2):
            got = _corrected_quotient(g, L, _multiplier(L, ulps))
            bad = np.nonzero(got != want)[0]
            assert bad.size == 0, (ulps, g[bad[:3]], L[bad[:3]], got[bad[:3]], want[bad[:3]])
