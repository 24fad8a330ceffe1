"""Shared helpers of the test-suite: fixtures under tests/golden/ -> inputs of the provider and of the oracle."""
import os
from types import SimpleNamespace

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    return g


def basis_objects(g):
    """list of plain objects with the reference Basis attributes, holding the reference's own normalised values."""
    off = np.concatenate([[0], np.cumsum(g["nprim"])])
    out = []
    for i in range(int(g["ncart"])):
        s = slice(off[i], off[i + 1])
        out.append(SimpleNamespace(origin=g["origins"][i], shell=g["lmn"][i], num_exps=int(g["nprim"][i]),
                                   exps=g["exps"][s], coefs=g["coefs"][s], norm=g["norms"][s]))
    return out


def oracle_basis(oracle, g):
    return oracle.FlatBasis(g["origins"], g["lmn"], g["nprim"], g["exps"], g["coefs"], g["norms"])


def context_for(g, device=0):
    import tuna_b200
    from tuna_b200.basis import flatten
    ctx = tuna_b200.Context(device)
    ctx.set_basis(*flatten(basis_objects(g)))
    return ctx


def eri_tolerance(ref):
    """1e-12 Eh absolute, relaxed to 1e-13 relative where |ERI| > 10."""
    return np.maximum(1e-12, 1e-13 * np.abs(ref))


def mo_inputs(g, n=None, seed=20261018):
    """Deterministic inputs of the AO->MO transformation tests (the same on the dev box, where the reference produced
    tests/golden/mo_transform.npz from them, and on the GPU box).  With a golden config g: MO-like coefficients
    C = X Q (X = S^-1/2 from the fixture, Q a seeded orthogonal matrix), a second set C_beta and orbital energies for the
    spin-blocking of tuna_ci.py:111-132.  With g = None: a synthetic dense tensor E of dimension n as well."""
    rng = np.random.default_rng(seed)
    out = {}
    if g is None:
        out["E"] = rng.standard_normal((n, n, n, n))
        X = np.eye(n) + 0.1 * rng.standard_normal((n, n))
    else:
        n = int(g["nbf"])
        X = np.array(g["X"])
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    Qb, _ = np.linalg.qr(rng.standard_normal((n, n)))
    out["C"] = np.ascontiguousarray(X @ Q)
    out["C_beta"] = np.ascontiguousarray(X @ Qb)
    out["eps"] = rng.standard_normal(2 * n)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# full-size parity of direct J/K without a dense oracle tensor: unit-pair densities
# ---------------------------------------------------------------------------------------------------------------------
def pick_function(fb, atom, L, tight, comp=0):
    """Index of component `comp` of the tightest / most diffuse shell of angular momentum L on atom 0 / 1 of a diatomic FlatBasis."""
    z = fb.origins[:, 2]
    Ls = np.asarray(fb.lmn).sum(axis=1)
    m = np.where((Ls == L) & (z == (z.min() if atom == 0 else z.max())))[0]
    e = np.asarray(fb.exps)[np.asarray(fb.offsets)[m]]
    m = m[e == (e.max() if tight else e.min())]
    return int(m[comp])


def unit_pair_density(n, k, l):
    """P = e_k e_l^T + e_l e_k^T: then J_ij = (ij|kl) + (ij|lk) and K_ij = (il|kj) + (ik|lj) (tuna_scf.py:42,70)."""
    P = np.zeros((n, n))
    P[k, l] += 1.0
    P[l, k] += 1.0
    return P


def check_unit_pair_jk(oracle, fb, k, l, J, K, n_samples=300, seed=11):
    """Compare sampled elements of J and K built from unit_pair_density(k, l) with single integrals of the oracle.
    Tolerance: two ERIs, each within SURVEY.md 8(d)'s sweep tolerance max(1e-12, 1e-13 |ERI|).  Returns the worst abs error."""
    rng = np.random.default_rng(seed)
    n = fb.ncart
    worst = 0.0
    for i, j in zip(rng.integers(0, n, n_samples), rng.integers(0, n, n_samples)):
        i, j = int(i), int(j)
        a, b = oracle.eri_single(fb, i, j, k, l), oracle.eri_single(fb, i, j, l, k)
        c, d = oracle.eri_single(fb, i, l, k, j), oracle.eri_single(fb, i, k, l, j)
        for got, ref, scale in ((J[i, j], a + b, max(abs(a), abs(b))), (K[i, j], c + d, max(abs(c), abs(d)))):
            err = abs(got - ref)
            assert err <= 2 * max(1e-12, 1e-13 * scale), f"({i},{j}|{k},{l}): got {got!r}, oracle {ref!r}, diff {err:.2e}"
            worst = max(worst, err)
    return worst
