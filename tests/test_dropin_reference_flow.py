"""The drop-in boundary exercised by the UNMODIFIED reference program (dev container only; skipped where /root/reference is absent).

`tuna_b200.install()` rebinds the ten names of INTEGRATION.md section 2 on the reference's modules and the reference's own driver
(tuna_energy.evaluate_molecular_energy) runs whole calculations through the provider's Python layer: ERIHandle through
`Integrals.ERI_AO`, the J-then-K cache, the spherical transformation, one-electron integrals, the guess's cross-basis overlap, and
the post-HF AO->MO transformation.  There is no GPU here, so the CUDA library behind `tuna_b200._lib.Context` is replaced — IN THIS
TEST ONLY — by a test double that answers the same methods from the oracle (tests may use the oracle as the checker; the product has no
such path).  What is verified is therefore the glue, against the reference's own numbers: energies and SCF iteration counts."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import ref_harness as rh  # noqa: E402
from util import load_golden  # noqa: E402

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference tree not present (GPU box)")


class OracleBackedContext:
    """Test double for tuna_b200._lib.Context: same methods and conventions, answered by the oracle."""

    def __init__(self, device=0):
        from oracle import tuna_oracle
        self.orc, self.device = tuna_oracle, device
        self.ncart = self.nbf = self.n_stored = 0
        self.fb = self.U = self.E_cart = self.E_sph = None

    def set_basis(self, oz, lmn, nprim, exps, ceff):
        oz = np.asarray(oz, dtype=np.float64)
        origins = np.stack([np.zeros_like(oz), np.zeros_like(oz), oz], axis=1)
        ceff = np.asarray(ceff, dtype=np.float64)
        self.fb = self.orc.FlatBasis(origins, np.asarray(lmn).reshape(-1, 3), np.asarray(nprim), np.asarray(exps, dtype=np.float64), ceff, np.ones_like(ceff))
        self.ncart, self.nbf, self.n_stored, self.E_cart, self.E_sph = len(oz), 0, 0, None, None

    def set_transform(self, U):
        self.U = np.array(U, dtype=np.float64)
        self.nbf = self.U.shape[0]

    def eri_fill_cart(self):
        self.E_cart = self.orc.eri_fill(self.fb)

    def eri_cart_to_sph(self, keep_cart=False):
        self.E_sph = self.orc.cart_to_sph_eri(self.E_cart, self.U)
        self.n_stored = self.nbf

    def _sph(self):
        if self.E_sph is None:
            if self.E_cart is None:
                self.eri_fill_cart()
            self.E_sph = self.orc.cart_to_sph_eri(self.E_cart, self.U)
        return self.E_sph

    def eri_download(self, which, out=None):
        src = self.E_cart if which == 0 else self.E_sph
        if out is None:
            return src.copy()
        out[...] = src
        return out

    def eri_upload(self, tensor):
        self.E_sph = np.array(tensor, dtype=np.float64)
        self.n_stored = self.E_sph.shape[0]

    def eri_single(self, i, j, k, l):
        return self.orc.eri_single(self.fb, i, j, k, l)

    def _jk(self, P, E, want_j, want_k):
        P = np.asarray(P, dtype=np.float64)
        Ps = P[None] if P.ndim == 2 else P
        J = np.stack([self.orc.coulomb(p, E) for p in Ps]) if want_j else None
        K = np.stack([self.orc.exchange(p, E) for p in Ps]) if want_k else None
        if P.ndim == 2:
            return (J[0] if want_j else None), (K[0] if want_k else None)
        return J, K

    def jk_stored(self, P, want_j=True, want_k=True):
        return self._jk(P, self.E_sph, want_j, want_k)

    def jk_direct(self, P, tau=1e-16, want_j=True, want_k=True):
        return self._jk(P, self._sph(), want_j, want_k)

    def eri_transform(self, C1, C2=None, so_layout=False, eri=None):
        E = self.E_sph if eri is None else np.asarray(eri)
        C2 = C1 if C2 is None else C2
        T = self.orc.transform_eri_ao_to_so(E, C1, C2)
        return T if so_layout else np.ascontiguousarray(np.einsum("pqrs->prqs", T))

    def eri_transform_spin_blocked(self, C1, C2=None, so_layout=True):
        G = np.kron(np.eye(2), np.kron(np.eye(2), self.E_sph).T)          # tuna_ci.py:564
        return self.eri_transform(C1, C2, so_layout, eri=G)

    def one_electron(self, atom_z, atom_charge, dipole_origin):
        return self.orc.one_electron(self.fb, atom_z, atom_charge, dipole_origin)

    def cross_overlap(self, f1, f2):
        def fb(f):
            oz, lmn, nprim, exps, ce = f
            oz = np.asarray(oz, dtype=np.float64)
            return self.orc.FlatBasis(np.stack([np.zeros_like(oz), np.zeros_like(oz), oz], axis=1), np.asarray(lmn).reshape(-1, 3), np.asarray(nprim),
                                      np.asarray(exps, dtype=np.float64), np.asarray(ce, dtype=np.float64), np.ones(len(ce)))
        return self.orc.cross_overlap(fb(f1), fb(f2))

    def close(self):
        pass


@pytest.fixture()
def installed(monkeypatch):
    import tuna_b200
    from tuna_b200 import _lib, provider
    ns = rh.load_reference()
    __import__("tuna_ci")
    monkeypatch.setattr(_lib, "Context", OracleBackedContext)
    monkeypatch.setattr(provider, "_free_device_bytes", lambda: 64 * 2 ** 30)
    monkeypatch.setattr(provider, "_scratch", {})
    saved_err = _lib.error_class
    originals = tuna_b200.install()
    try:
        yield ns, tuna_b200
    finally:
        tuna_b200.uninstall(originals)
        _lib.error_class = saved_err
        tuna_b200.configure(mode="auto")


def run(ns, line):
    calc, symbols, coords = rh.parse_line(ns, line)
    out, molecule, energy, P = ns.energ.evaluate_molecular_energy(calc, symbols, coords, silent=True)
    return float(energy), out


@pytest.mark.parametrize("mode", ["stored", "direct"])
def test_h2_scf_through_the_installed_provider(installed, mode):
    ns, tb = installed
    tb.configure(mode=mode)
    g = load_golden("h2_631g_nodiis")
    energy, out = run(ns, str(g["line"]))
    assert abs(energy - float(g["energy"])) < 1e-10
    assert abs(float(out.coulomb_energy) - float(g["coulomb_energy"])) < 1e-10 and abs(float(out.exchange_energy) - float(g["exchange_energy"])) < 1e-10


def test_n2_ccpvtz_scf_through_the_installed_provider(installed):
    ns, tb = installed
    tb.configure(mode="auto")
    g = load_golden("n2_ccpvtz")
    energy, out = run(ns, str(g["line"]))
    assert abs(energy - float(g["energy"])) < 1e-10          # -108.983006526214


def test_post_hf_consumers_get_what_they_need(installed):
    """MP2 on H2/6-31G: `auto` keeps the dense tensor, the AO->MO / spin-orbital transformations go through the provider, and the total
    energy equals the unpatched reference's."""
    ns, tb = installed
    names = {m.name for m in ns.util.electronic_structure_methods}
    if "MP2" not in names:
        pytest.skip("reference build without MP2")
    line = "SPE : H H 0.74 : MP2 6-31G : NODIIS"
    calls = {"mo": 0, "so": 0}
    import tuna_ci
    mo, so = tuna_ci.transform_ERI_AO_to_MO, tuna_ci.transform_ERI_AO_to_SO
    tuna_ci.transform_ERI_AO_to_MO = lambda *a, **k: (calls.__setitem__("mo", calls["mo"] + 1), mo(*a, **k))[1]
    tuna_ci.transform_ERI_AO_to_SO = lambda *a, **k: (calls.__setitem__("so", calls["so"] + 1), so(*a, **k))[1]
    try:
        e_provider, _ = run(ns, line)
    finally:
        tuna_ci.transform_ERI_AO_to_MO, tuna_ci.transform_ERI_AO_to_SO = mo, so
    assert calls["mo"] + calls["so"] >= 1
    assert abs(e_provider - run_unpatched(ns, line)) < 1e-10


def run_unpatched(ns, line):
    """The same line with the reference's own ten functions temporarily restored."""
    saved = {key: getattr(sys.modules[key[0]], key[1]) for key in _REF_ORIGINALS}
    try:
        for (m, n), fn in _REF_ORIGINALS.items():
            setattr(sys.modules[m], n, fn)
        return run(ns, line)[0]
    finally:
        for (m, n), fn in saved.items():
            setattr(sys.modules[m], n, fn)


@pytest.mark.parametrize("line", ["SPE : H H 0.74 : UHF 6-31G : NOROTATE NODIIS", "SPE : H H 0.74 : HF 6-31G : CARTHARM NODIIS", "SPE : HE : HF 6-31G : NODIIS",
                                  "SPE : H H 0.74 : B3LYP 6-31G : NODIIS", "SPE : H H 0.74 : CCSD 6-31G : NODIIS", "SPE : H H 0.74 : CIS 6-31G : NODIIS",
                                  "SPE : H H 0.74 : UMP2 6-31G : NOROTATE NODIIS", "SPE : H H 0.74 : B2PLYP 6-31G : NODIIS"])
def test_other_flows_match_the_unpatched_reference(installed, line):
    """UHF (J/K per spin), CARTHARM (no rotation, tuna_kernel.py:481-483), a single atom, hybrid DFT (exact exchange through K), and the
    post-HF consumers of the dense tensor: coupled cluster, CIS, unrestricted MP2 (three spin-orbital transformations, tuna_mp.py:1055-1057),
    a double hybrid.
    NODIIS: with default keywords H2/6-31G is round-off chaotic in the reference itself (1e-15 perturbations move E by 1e-7)."""
    ns, tb = installed
    tb.configure(mode="auto")
    try:
        e_provider, _ = run(ns, line)
    except StopIteration:
        pytest.skip("method not in this reference build")
    assert abs(e_provider - run_unpatched(ns, line)) < 1e-10


_REF_ORIGINALS = {}


@pytest.fixture(autouse=True)
def _remember_reference_functions():
    """Snapshot of the reference's own ten functions, taken before any test installs the provider."""
    if not _REF_ORIGINALS:
        ns = rh.load_reference()
        ci = __import__("tuna_ci")
        for mod, names in ((ns.ints, ("calculate_electron_repulsion_integrals", "calculate_electron_repulsion_integral", "calculate_one_electron_integrals",
                                      "calculate_cross_basis_overlap_matrix")),
                           (ns.kern, ("calculate_two_electron_integrals", "transform_to_spherical_harmonics")),
                           (ns.scf, ("calculate_coulomb_matrix", "calculate_exchange_matrix")),
                           (ci, ("transform_ERI_AO_to_MO", "transform_ERI_AO_to_SO"))):
            for n in names:
                _REF_ORIGINALS[(mod.__name__, n)] = getattr(mod, n)
    yield


@pytest.mark.parametrize("mode", ["stored", "direct"])
def test_general_density_through_the_reference_signatures(installed, mode):
    """The reference's J/K accept ANY real matrix (DIIS / damped / guess densities are not exactly symmetric): so do the provider's, in both modes."""
    ns, tb = installed
    from types import SimpleNamespace
    from oracle import tuna_oracle as orc
    from util import basis_objects, oracle_basis
    g = load_golden("h2_631g")
    bfs = basis_objects(g)
    tb.configure(mode=mode)
    calc = SimpleNamespace(cartesian_harmonics=False, number_of_threads=1, method=SimpleNamespace(method_base="HF"))
    h = tb.calculate_two_electron_integrals(len(bfs), bfs, calc)
    one = np.eye(len(bfs))
    *_, eri = tb.transform_to_spherical_harmonics(one, one, one, np.stack([one] * 3), np.stack([one] * 3), h, SimpleNamespace(spherical_harmonic_transformation_matrix=g["U"]), calc, True)
    assert eri.mode == mode
    P = np.random.default_rng(2).standard_normal((int(g["nbf"]),) * 2)
    E = orc.cart_to_sph_eri(orc.eri_fill(oracle_basis(orc, g)), np.array(g["U"]))
    assert np.abs(tb.calculate_coulomb_matrix(P, eri) - np.einsum("ijkl,kl->ij", E, P)).max() < 1e-12
    assert np.abs(tb.calculate_exchange_matrix(P, eri) - np.einsum("ilkj,kl->ij", E, P)).max() < 1e-12
