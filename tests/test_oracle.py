"""Pin the oracle: C restatement vs the reference's compiled engine (oracle/_ref), the KATs of SURVEY.md 8(c)
and the committed golden fixtures.  CPU only."""
import os

import numpy as np
import pytest

from util import load_golden, oracle_basis

# unique ERIs of H2/6-31G at 0.74 Angstrom harvested from the unmodified reference
H2_KAT = {(0, 0, 0, 0): 1.076566132473073, (1, 0, 0, 0): 0.578470297074810, (1, 0, 1, 0): 0.329423122075922,
          (1, 1, 0, 0): 0.587395831106644, (1, 1, 1, 0): 0.361303768639465, (1, 1, 1, 1): 0.453150328467739,
          (2, 0, 0, 0): 0.408221754767887, (2, 0, 1, 0): 0.232157469110124, (2, 0, 1, 1): 0.252976506391250,
          (2, 0, 2, 0): 0.200338809231908, (2, 1, 0, 0): 0.346642955290003, (2, 1, 1, 0): 0.209953976151585,
          (2, 1, 1, 1): 0.252683167942554, (2, 1, 2, 0): 0.187985219541137, (2, 1, 2, 1): 0.196048647372463,
          (2, 2, 0, 0): 0.662261550005863, (3, 1, 1, 1): 0.377106075661827, (3, 1, 3, 1): 0.330578063140490,
          (3, 3, 3, 3): 0.453150328467739}


def test_ss_ss_unit_exponent(oracle):
    fb = oracle.FlatBasis.from_raw(np.zeros((1, 3)), np.zeros((1, 3), dtype=int), [1], [1.0], [1.0])
    assert abs(oracle.eri_single(fb, 0, 0, 0, 0) - 1.128379167095513) < 1e-15     # 2/sqrt(pi)


def test_h2_known_answers(oracle):
    g = load_golden("h2_631g")
    E = oracle.eri_fill(oracle_basis(oracle, g))
    for idx, val in H2_KAT.items():
        assert abs(E[idx] - val) < 2e-15, idx
    assert abs(E.sum() - float(g["eri_cart_sum"])) < 1e-12


def test_normalisation_matches_reference_values(oracle):
    g = load_golden("n2_ccpvtz")
    fb = oracle.FlatBasis.from_raw(g["origins"], g["lmn"], g["nprim"], g["exps"], g["coefs"])   # re-normalising is idempotent
    np.testing.assert_allclose(fb.coefs, g["coefs"], rtol=1e-13)
    np.testing.assert_allclose(fb.norms, g["norms"], rtol=1e-14)
    # H2/6-31G normalised values quoted in SURVEY.md 8(c)
    h = load_golden("h2_631g")
    np.testing.assert_allclose(h["coefs"][:3], [0.033494604341276, 0.234726953508943, 0.813757326131003], rtol=1e-13)
    np.testing.assert_allclose(h["norms"][:4], [6.417017102526964, 1.553171447742897, 0.51004324571929, 0.181380649178652], rtol=1e-13)


@pytest.mark.parametrize("name", ["n2_ccpvtz", "et100"])
def test_oracle_vs_golden_samples(oracle, name):
    g = load_golden(name)
    E = oracle.eri_fill(oracle_basis(oracle, g))
    ref = g["eri_cart_val"]
    got = E[tuple(g["eri_cart_idx"].T)]
    assert np.all(np.abs(got - ref) <= np.maximum(1e-12, 1e-13 * np.abs(ref)))
    assert abs(E.sum() - float(g["eri_cart_sum"])) < 1e-8 * abs(float(g["eri_cart_sum"]))
    assert abs(np.linalg.norm(E) - float(g["eri_cart_fro"])) < 1e-11 * float(g["eri_cart_fro"])
    # rotation + J/K restatements against the reference's own outputs
    Es = oracle.cart_to_sph_eri(E, g["U"])
    np.testing.assert_allclose(Es[tuple(g["eri_sph_idx"].T)], g["eri_sph_val"], atol=2e-12, rtol=0)
    from tuna_b200.workloads import fixed_density
    P = fixed_density(int(g["nbf"]))
    np.testing.assert_allclose(oracle.coulomb(P, Es), g["Jfix"], atol=1e-10, rtol=0)
    np.testing.assert_allclose(oracle.exchange(P, Es), g["Kfix"], atol=1e-10, rtol=0)


def test_n2_ccpvtz_published_checksums(oracle):
    """J/K checksums on the fixed density for N2/cc-pVTZ."""
    g = load_golden("n2_ccpvtz")
    assert abs(g["Jfix"].sum() - 9.4946892380) < 1e-8 and abs(np.linalg.norm(g["Jfix"]) - 36.5698933327) < 1e-8
    assert abs(g["Kfix"].sum() - 245.5653859429) < 1e-8 and abs(np.linalg.norm(g["Kfix"]) - 54.3720356773) < 1e-8
    assert abs(g["Jfix"][0, 0] - 13.037335815535) < 1e-10 and abs(g["Kfix"][0, 0] - 13.518914795950) < 1e-10


def test_oracle_vs_compiled_reference(oracle):
    """Element-wise against the UNMODIFIED reference engine (oracle/_ref) on the full N2/cc-pVTZ Cartesian tensor."""
    eng = oracle.reference_engine()
    if eng is None:
        pytest.skip("oracle/_ref not built (needs /root/reference once; the .so then travels with the repo)")
    g = load_golden("n2_ccpvtz")
    fb = oracle_basis(oracle, g)
    n = fb.ncart
    ref = np.asarray(eng.calculate_electron_repulsion_integrals(n, np.empty((n,) * 4), oracle.reference_basis_objects(fb), 8))
    got = oracle.eri_fill(fb)
    assert np.abs(got - ref).max() < 1e-13
    # the parity zeros are exact zeros in both
    assert np.array_equal(got == 0.0, ref == 0.0)
    # single-quartet entry point (pyx:1376-1414), including an x-parity zero
    bfs = oracle.reference_basis_objects(fb)
    for q in [(0, 0, 0, 0), (7, 40, 1, 41), (17, 60, 18, 58), (30, 69, 22, 50)]:
        r = eng.calculate_electron_repulsion_integral(*[bfs[i] for i in q])
        assert abs(oracle.eri_single(fb, *q) - r) < 1e-13


def test_boys_against_series(oracle):
    """Top-order Boys function vs an independent high-precision evaluation (mpmath-free: long Kummer series in Python)."""
    from fractions import Fraction
    import math
    for m in (0, 3, 8, 20):
        for T in (0.0, 1e-3, 0.7, 5.0, 18.0, 34.9, 35.1, 60.0, 500.0):
            if T < 30:
                term, s = 1.0 / (2 * m + 1), 0.0
                terms = []
                for k in range(1, 600):
                    terms.append(term)
                    term *= 2 * T / (2 * m + 2 * k + 1)
                s = math.fsum(terms) * math.exp(-T)
            else:
                # asymptotic + exact upward recursion in high precision
                from decimal import Decimal, getcontext
                getcontext().prec = 50
                Td = Decimal(T)
                e = (-Td).exp()
                F = Decimal(math.pi).sqrt() / (2 * Td.sqrt()) * Decimal(math.erf(math.sqrt(T)))
                for k in range(m):
                    F = ((2 * k + 1) * F - e) / (2 * Td)
                s = float(F)
            assert abs(oracle.boys(m, T) - s) <= 5e-15 * abs(s) + 1e-300, (m, T)
