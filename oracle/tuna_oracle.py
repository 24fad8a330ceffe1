"""CPU oracle for the TUNA SCF two-electron hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this
module, and only as the checker.  The product package (tuna_b200/) never does.  Parity status: PINNED against
the reference's own compiled engine (oracle/_ref) and the known-answer vectors of SURVEY.md section 8(c);
see tests/test_oracle.py.

Restated here in NumPy (file:line under /root/reference/TUNA/):
    coulomb / exchange      <- tuna_scf.py:55-72, :27-44     (the two einsums)
    cart_to_sph_eri         <- tuna_kernel.py:504-523         ((U (x) U) ERI (U (x) U)^T)
    transform_eri_ao_to_mo  <- tuna_ci.py:204-255             (four einsums, layout prqs)
    transform_eri_ao_to_so  <- tuna_ci.py:143-193, :564       (four einsums, layout pqrs; spin blocking)
    FlatBasis.from_*        <- tuna_integrals/tuna_integral.pyx:78-235 (Basis + normalize)
    eri_fill / eri_single   -> oracle/eri_oracle.c            (pyx:961-1414, :1490-1651)
    one_electron / cross_overlap -> oracle/oneel_oracle.c     (pyx:282-912, :1428-1489)
"""
import ctypes
import os
import subprocess
import sys
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile oracle/eri_oracle.c (gcc) and, if the reference tree is here, oracle/_ref."""
    subprocess.run(["make", "-s", "-C", _HERE, "liboracle.so"], check=True)
    subprocess.run(["bash", os.path.join(_HERE, "build_ref.sh")], check=True, stdout=subprocess.DEVNULL)


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        srcs = [os.path.join(_HERE, f) for f in ("eri_oracle.c", "oneel_oracle.c")]
        if not os.path.exists(path) or any(os.path.getmtime(f) > os.path.getmtime(path) for f in srcs):
            subprocess.run(["make", "-s", "-C", _HERE, "liboracle.so"], check=True)
        lib = ctypes.CDLL(path)
        dp, ip, lp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_long)
        lib.oracle_eri_fill.argtypes = [ctypes.c_long, dp, ip, lp, lp, dp, dp, dp, dp, ctypes.c_int]
        lib.oracle_eri_fill.restype = ctypes.c_int
        lib.oracle_eri_single.argtypes = [dp, ip, lp, lp, dp, dp, dp] + [ctypes.c_long] * 4
        lib.oracle_eri_single.restype = ctypes.c_double
        lib.oracle_boys.argtypes = [ctypes.c_int, ctypes.c_double]
        lib.oracle_boys.restype = ctypes.c_double
        lib.oracle_normalize.argtypes = [ctypes.c_int] * 3 + [ctypes.c_long, dp, dp, dp]
        lib.oracle_normalize.restype = None
        lib.oracle_max_threads.restype = ctypes.c_int
        lib.oracle_one_electron.argtypes = [ctypes.c_long, dp, ip, lp, lp, dp, dp, ctypes.c_long, dp, dp, dp, dp, dp, dp, dp, dp, ctypes.c_int]
        lib.oracle_one_electron.restype = ctypes.c_int
        lib.oracle_cross_overlap.argtypes = [ctypes.c_long, dp, ip, lp, lp, dp, dp] * 2 + [dp]
        lib.oracle_cross_overlap.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


@dataclass
class FlatBasis:
    """The list[Basis] of the reference (one entry per CARTESIAN COMPONENT, pyx:78-235) as flat arrays."""
    origins: np.ndarray   # (ncart, 3)
    lmn: np.ndarray       # (ncart, 3) int
    nprim: np.ndarray     # (ncart,)
    exps: np.ndarray      # concatenated
    coefs: np.ndarray     # contraction-normalised coefficients (what Basis.coefs holds after normalize)
    norms: np.ndarray     # primitive norms (Basis.norm)

    @property
    def ncart(self):
        return len(self.nprim)

    @property
    def offsets(self):
        return np.concatenate([[0], np.cumsum(self.nprim)[:-1]]).astype(np.int64)

    @classmethod
    def from_raw(cls, origins, lmn, nprim, exps, raw_coefs):
        """Build from un-normalised contraction coefficients, applying Basis.normalize (pyx:174-210)."""
        origins = np.ascontiguousarray(origins, dtype=np.float64).reshape(-1, 3)
        lmn = np.ascontiguousarray(lmn, dtype=np.int64).reshape(-1, 3)
        nprim = np.ascontiguousarray(nprim, dtype=np.int64)
        exps = np.ascontiguousarray(exps, dtype=np.float64)
        coefs = np.array(raw_coefs, dtype=np.float64)
        norms = np.zeros_like(coefs)
        lib, off = _lib(), 0
        for i, n in enumerate(nprim):
            e, c, nn = exps[off:off + n].copy(), coefs[off:off + n].copy(), np.zeros(n)
            lib.oracle_normalize(int(lmn[i, 0]), int(lmn[i, 1]), int(lmn[i, 2]), int(n),
                                 _p(e, ctypes.c_double), _p(c, ctypes.c_double), _p(nn, ctypes.c_double))
            coefs[off:off + n], norms[off:off + n] = c, nn
            off += n
        return cls(origins, lmn, nprim, exps, coefs, norms)

    @classmethod
    def from_reference_objects(cls, bfs):
        """From a list of reference (or provider) Basis objects: reads .origin/.shell/.num_exps/.exps/.coefs/.norm."""
        n = len(bfs)
        return cls(np.array([np.array(b.origin) for b in bfs], dtype=np.float64).reshape(n, 3),
                   np.array([np.array(b.shell) for b in bfs], dtype=np.int64).reshape(n, 3),
                   np.array([int(b.num_exps) for b in bfs], dtype=np.int64),
                   np.concatenate([np.array(b.exps, dtype=np.float64) for b in bfs]),
                   np.concatenate([np.array(b.coefs, dtype=np.float64) for b in bfs]),
                   np.concatenate([np.array(b.norm, dtype=np.float64) for b in bfs]))

    def _c_args(self):
        self._keep = (np.ascontiguousarray(self.origins[:, 2]), np.ascontiguousarray(self.lmn, dtype=np.int32),
                      np.ascontiguousarray(self.nprim, dtype=np.int64), np.ascontiguousarray(self.offsets, dtype=np.int64),
                      np.ascontiguousarray(self.exps), np.ascontiguousarray(self.coefs), np.ascontiguousarray(self.norms))
        oz, lmn, npr, off, ex, co, no = self._keep
        return (_p(oz, ctypes.c_double), _p(lmn, ctypes.c_int), _p(npr, ctypes.c_long), _p(off, ctypes.c_long),
                _p(ex, ctypes.c_double), _p(co, ctypes.c_double), _p(no, ctypes.c_double))


def max_threads():
    return int(_lib().oracle_max_threads())


def eri_fill(basis: FlatBasis, nthreads: int = 0) -> np.ndarray:
    """Dense Cartesian ERI tensor (ncart^4, C order), restating pyx:1267-1355."""
    n = basis.ncart
    out = np.empty((n, n, n, n))
    args = basis._c_args()
    rc = _lib().oracle_eri_fill(n, *args, _p(out, ctypes.c_double), nthreads or max_threads())
    if rc:
        raise MemoryError()
    return out


def eri_single(basis: FlatBasis, i, j, k, l) -> float:
    return float(_lib().oracle_eri_single(*basis._c_args(), i, j, k, l))


def _flat_args(basis: FlatBasis):
    """(oz, lmn, nprim, offsets, exps, coef_eff) for the one-electron entry points; coef_eff = norm * coefs (pyx:504-508)."""
    keep = (np.ascontiguousarray(basis.origins[:, 2]), np.ascontiguousarray(basis.lmn, dtype=np.int32), np.ascontiguousarray(basis.nprim, dtype=np.int64),
            np.ascontiguousarray(basis.offsets, dtype=np.int64), np.ascontiguousarray(basis.exps), np.ascontiguousarray(basis.coefs * basis.norms))
    return keep, (_p(keep[0], ctypes.c_double), _p(keep[1], ctypes.c_int), _p(keep[2], ctypes.c_long), _p(keep[3], ctypes.c_long),
                  _p(keep[4], ctypes.c_double), _p(keep[5], ctypes.c_double))


def one_electron(basis: FlatBasis, atom_z, atom_charge, dipole_origin):
    """(S, T, V_NE, D[3], Q[3]) in the Cartesian basis, restating calculate_one_electron_integrals (pyx:282-445)."""
    n = basis.ncart
    az = np.ascontiguousarray(atom_z, dtype=np.float64)
    ac = np.ascontiguousarray(atom_charge, dtype=np.float64)
    og = np.ascontiguousarray(dipole_origin, dtype=np.float64)
    S, T, V, D, Q = np.empty((n, n)), np.empty((n, n)), np.empty((n, n)), np.empty((3, n, n)), np.empty((3, n, n))
    keep, args = _flat_args(basis)
    d = ctypes.c_double
    _lib().oracle_one_electron(n, *args, len(az), _p(az, d), _p(ac, d), _p(og, d), _p(S, d), _p(T, d), _p(V, d), _p(D, d), _p(Q, d), 0)
    return S, T, V, D, Q


def cross_overlap(basis_1: FlatBasis, basis_2: FlatBasis):
    """S12[i, j] = <bf_1[i] | bf_2[j]>, restating calculate_cross_basis_overlap_matrix (pyx:626-778)."""
    out = np.empty((basis_1.ncart, basis_2.ncart))
    k1, a1 = _flat_args(basis_1)
    k2, a2 = _flat_args(basis_2)
    _lib().oracle_cross_overlap(basis_1.ncart, *a1, basis_2.ncart, *a2, _p(out, ctypes.c_double))
    return out


def boys(m: int, T: float) -> float:
    return float(_lib().oracle_boys(m, T))


def coulomb(P, ERI):      # tuna_scf.py:70
    return np.einsum("ijkl,kl->ij", ERI, P, optimize=True)


def exchange(P, ERI):     # tuna_scf.py:42
    return np.einsum("ilkj,kl->ij", ERI, P, optimize=True)


def cart_to_sph_eri(ERI_cart, U):
    """ERI_sph = (U(x)U) ERI_cart (U(x)U)^T, tuna_kernel.py:504-523, as four dense index rotations."""
    t = np.einsum("pi,ijkl->pjkl", U, ERI_cart, optimize=True)
    t = np.einsum("qj,pjkl->pqkl", U, t, optimize=True)
    t = np.einsum("rk,pqkl->pqrl", U, t, optimize=True)
    return np.einsum("sl,pqrl->pqrs", U, t, optimize=True)


def transform_eri_ao_to_mo(ERI_AO, C):
    """tuna_ci.py:204-255: the four stepwise einsums of transform_ERI_AO_to_MO; result in interleaved chemists' layout [p][r][q][s]."""
    t = np.einsum("mknl,ls->mnks", ERI_AO, C, optimize=True)      # :230
    t = np.einsum("mnks,kr->mnrs", t, C, optimize=True)           # :236
    t = np.einsum("mnrs,nq->mqrs", t, C, optimize=True)           # :242
    return np.einsum("mqrs,mp->prqs", t, C, optimize=True)        # :250


def transform_eri_ao_to_so(ERI_AO, C_1, C_2):
    """tuna_ci.py:143-193: transform_ERI_AO_to_SO; result in physicists' layout [p][q][r][s]."""
    t = np.einsum("mknl,ls->mnks", ERI_AO, C_1, optimize=True)    # :165
    t = np.einsum("mnks,kr->mnrs", t, C_2, optimize=True)         # :171
    t = np.einsum("mnrs,nq->mqrs", t, C_1, optimize=True)         # :177
    return np.einsum("mqrs,mp->pqrs", t, C_2, optimize=True)      # :185


def spin_block_eri(ERI_AO):
    """tuna_ci.py:564: ERI_spin_block = kron(I2, kron(I2, ERI_AO).T)."""
    return np.kron(np.eye(2), np.kron(np.eye(2), ERI_AO).T)


def parity_surviving_quartets(lmn) -> tuple:
    """(unique AO quartets, those passing the x/y parity test of pyx:1324-1327) — the 'ERI quartets' unit."""
    lmn = np.asarray(lmn)
    n = len(lmn)
    i, j = np.tril_indices(n)
    px = (lmn[i, 0] + lmn[j, 0]) & 1
    py = (lmn[i, 1] + lmn[j, 1]) & 1
    cls = px * 2 + py
    counts = np.bincount(cls, minlength=4).astype(np.int64)
    npair = len(i)
    unique = npair * (npair + 1) // 2
    surviving = int(sum(c * (c + 1) // 2 for c in counts))
    return unique, surviving


def reference_engine():
    """The UNMODIFIED reference engine compiled into oracle/_ref by build_ref.sh (None if not built)."""
    d = os.path.join(_HERE, "_ref")
    if d not in sys.path:
        sys.path.insert(0, d)
    try:
        import tuna_integral
        return tuna_integral
    except ImportError:
        return None


def reference_basis_objects(basis: FlatBasis, raw_coefs=None):
    """Instantiate the reference's own Basis objects (pyx:144-170) from raw (un-normalised) coefficients."""
    eng = reference_engine()
    off = basis.offsets
    out = []
    for i in range(basis.ncart):
        s = slice(off[i], off[i] + basis.nprim[i])
        c = raw_coefs[s] if raw_coefs is not None else basis.coefs[s]
        out.append(eng.Basis(basis.origins[i].copy(), basis.lmn[i].astype(np.int64), int(basis.nprim[i]),
                             basis.exps[s].copy(), np.array(c, dtype=np.float64)))
    return out
