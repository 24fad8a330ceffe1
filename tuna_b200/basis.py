"""Host-side mirror of the reference's `Basis` object and the flattening the C ABI needs.

`Basis` keeps the constructor signature and the read-only properties of the reference's cdef class
(TUNA/tuna_integrals/tuna_integral.pyx:78-235): one object per CARTESIAN COMPONENT of a contracted Gaussian,
normalised on construction exactly as `Basis.normalize` does (pyx:174-210).  The provider accepts either
these objects or the reference's own (anything exposing origin/shell/num_exps/exps/coefs/norm).
"""
from math import pi, sqrt

import numpy as np


def _dfact(n: int) -> float:
    r = 1.0
    while n > 1:
        r *= n
        n -= 2
    return r


class Basis:
    """Basis(origin, shell, num_exps, exps, coefs) — pyx:144-170."""

    __slots__ = ("origin", "shell", "num_exps", "exps", "coefs", "norm")

    def __init__(self, origin, shell, num_exps, exps, coefs):
        self.origin = np.array(origin, dtype=np.float64).reshape(3)
        self.shell = np.array(shell, dtype=np.int64).reshape(3)
        self.num_exps = int(num_exps)
        self.exps = np.array(exps, dtype=np.float64).reshape(self.num_exps)
        self.coefs = np.array(coefs, dtype=np.float64).reshape(self.num_exps)
        self.norm = np.zeros(self.num_exps)
        self.normalize()

    def normalize(self):
        """Primitive norms, then the contraction normalisation folded into coefs (pyx:174-210)."""
        l, m, n = (int(x) for x in self.shell)
        L = l + m + n
        df = _dfact(2 * l - 1) * _dfact(2 * m - 1) * _dfact(2 * n - 1)
        self.norm = np.sqrt(2.0 ** (2 * L + 1.5) * self.exps ** (L + 1.5) / df / pi ** 1.5)
        prefactor = pi ** 1.5 * df / 2.0 ** L
        w = self.norm * self.coefs
        N = float(np.sum(np.outer(w, w) / np.add.outer(self.exps, self.exps) ** (L + 1.5)))
        self.coefs = self.coefs / sqrt(prefactor * N)


def flatten(bfs):
    """list[Basis] -> (origins_z, lmn, nprim, exps, coef_eff) with coef_eff = norm * coefs (pyx:1070).

    Raises ValueError if a centre is off the z axis: the whole engine (like the reference's, tuna_kernel.py:386-388)
    assumes atoms and diatomics aligned on z.
    """
    n = len(bfs)
    origins = np.array([np.asarray(b.origin, dtype=np.float64) for b in bfs]).reshape(n, 3)
    if np.any(origins[:, :2] != 0.0):
        raise ValueError("all basis-function centres must lie on the z axis")
    lmn = np.array([np.asarray(b.shell) for b in bfs], dtype=np.int32).reshape(n, 3)
    nprim = np.array([int(b.num_exps) for b in bfs], dtype=np.int32)
    exps = np.concatenate([np.asarray(b.exps, dtype=np.float64) for b in bfs])
    ceff = np.concatenate([np.asarray(b.norm, dtype=np.float64) * np.asarray(b.coefs, dtype=np.float64) for b in bfs])
    return np.ascontiguousarray(origins[:, 2]), lmn, nprim, exps, ceff


def from_arrays(origins, lmn, nprim, exps, raw_coefs):
    """Build the list[Basis] from flat per-component arrays (e.g. tuna_b200.workloads.even_tempered_diatomic)."""
    out, off = [], 0
    for i, k in enumerate(nprim):
        k = int(k)
        out.append(Basis(origins[i], lmn[i], k, exps[off:off + k], raw_coefs[off:off + k]))
        off += k
    return out
