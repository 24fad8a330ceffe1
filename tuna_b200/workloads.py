"""Synthetic workloads named by BASELINE.json: the even-tempered diatomic sweep and the fixed test density.

Definitions follow SURVEY.md section 8(d), config 5: N2 at 1.10 Angstrom, fully uncontracted shells
alpha_{L,k} = alpha_L * beta^k with alpha_S..H = 0.10, 0.12, 0.30, 0.50, 0.80, 1.00 and per-atom compositions
    nbf100 = 10s6p3d1f (beta 2.5)      nbf200 = 13s9p6d3f1g (beta 2.5)
    nbf400 = 19s14p10d6f4g1h (beta 2.0) nbf800 = 32s24p17d12f8g5h (beta 1.6)
"""
import numpy as np

BOHR_PER_ANGSTROM = 1.8897261259065457
ET_ALPHA0 = (0.10, 0.12, 0.30, 0.50, 0.80, 1.00)
ET_SETS = {
    100: ((10, 6, 3, 1), 2.5),
    200: ((13, 9, 6, 3, 1), 2.5),
    400: ((19, 14, 10, 6, 4, 1), 2.0),
    800: ((32, 24, 17, 12, 8, 5), 1.6),
}
SHELL_LETTERS = "SPDFGH"


def cartesian_components(L: int):
    """Component order of the reference, TUNA/tuna_molecule.py:622: x^L ... z^L, i descending then j descending."""
    return [(i, j, L - i - j) for i in range(L, -1, -1) for j in range(L - i, -1, -1)]


def even_tempered_shells(nbf: int):
    """[(L, exponent)] for ONE atom of the named sweep point (each shell a single primitive, coefficient 1)."""
    counts, beta = ET_SETS[nbf]
    return [(L, ET_ALPHA0[L] * beta ** k) for L, n in enumerate(counts) for k in range(n)]


def shells_to_components(atom_shells, z_positions):
    """Expand per-atom shell lists [(L, exps, coefs)] into the reference's per-component basis arrays.

    Returns dict(origins, lmn, nprim, exps, raw_coefs) in the atom -> shell -> component order that
    tuna_molecule.form_basis produces (TUNA/tuna_molecule.py:532-585).
    """
    origins, lmn, nprim, exps, coefs = [], [], [], [], []
    for z, shells in zip(z_positions, atom_shells):
        for L, e, c in shells:
            for comp in cartesian_components(L):
                origins.append((0.0, 0.0, z))
                lmn.append(comp)
                nprim.append(len(e))
                exps.extend(e)
                coefs.extend(c)
    return dict(origins=np.array(origins, dtype=np.float64).reshape(-1, 3), lmn=np.array(lmn, dtype=np.int64).reshape(-1, 3),
                nprim=np.array(nprim, dtype=np.int64), exps=np.array(exps, dtype=np.float64),
                raw_coefs=np.array(coefs, dtype=np.float64))


def even_tempered_diatomic(nbf: int, bond_angstrom: float = 1.10):
    """The sweep point as raw per-component arrays (both atoms carry the same set)."""
    shells = [(L, [a], [1.0]) for L, a in even_tempered_shells(nbf)]
    return shells_to_components([shells, shells], [0.0, bond_angstrom * BOHR_PER_ANGSTROM])


def even_tempered_basis_file(nbf: int, element_name: str = "NITROGEN") -> str:
    """The same set in the reference's CUSTOM basis-file syntax (TUNA/tuna_basis.py:34-175)."""
    lines = [element_name]
    for L, a in even_tempered_shells(nbf):
        lines.append(f"{SHELL_LETTERS[L]} 1")
        lines.append(f" 1 {a!r} 1.0")
    return "\n".join(lines) + "\n"


def spherical_count(lmn) -> int:
    """nbf of the spherical basis for full Cartesian shells (2L+1 per (L+1)(L+2)/2 components)."""
    L = np.asarray(lmn).sum(axis=1)
    n = 0
    for l in range(int(L.max()) + 1):
        n += int((L == l).sum()) // ((l + 1) * (l + 2) // 2) * (2 * l + 1)
    return n


def fixed_density(n: int) -> np.ndarray:
    """P = (A + A^T)/2, A = default_rng(20261018).standard_normal((n, n)) — SURVEY.md section 8(d)."""
    A = np.random.default_rng(20261018).standard_normal((n, n))
    return (A + A.T) / 2
