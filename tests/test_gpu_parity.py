"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle, the reference's compiled engine
and the committed golden fixtures.  Tolerances are the north star's: ERIs 1e-12 Eh absolute (1e-13 relative above
10 Eh), J/K 1e-11, energies 1e-10 Eh."""
from types import SimpleNamespace

import numpy as np
import pytest

from test_oracle import H2_KAT
from util import basis_objects, context_for, eri_tolerance, load_golden, oracle_basis

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tb():
    import tuna_b200
    return tuna_b200


def _check_eri(got, ref):
    bad = np.abs(got - ref) > eri_tolerance(ref)
    assert not bad.any(), f"{int(bad.sum())} elements out of tolerance, max abs diff {np.abs(got - ref).max():.3e}"


def test_h2_known_answers(tb, oracle):
    g = load_golden("h2_631g")
    ctx = context_for(g)
    ctx.eri_fill_cart()
    E = ctx.eri_download(0)
    for idx, val in H2_KAT.items():
        assert abs(E[idx] - val) < 1e-12, idx
    _check_eri(E, oracle.eri_fill(oracle_basis(oracle, g)))
    for idx, val in list(H2_KAT.items())[:6]:
        assert abs(ctx.eri_single(*idx) - val) < 1e-12


@pytest.mark.parametrize("name", ["n2_ccpvtz", "co_b3lyp_ccpvtz", "et100"])
def test_eri_cart_vs_oracle(tb, oracle, name):
    """Dense Cartesian tensor (engine fill mode + k_fill_scatter) against the oracle and the reference's own recorded values."""
    g = load_golden(name)
    ctx = context_for(g)
    ctx.eri_fill_cart()
    E = ctx.eri_download(0)
    ref = oracle.eri_fill(oracle_basis(oracle, g))
    _check_eri(E, ref)
    assert np.array_equal(E == 0.0, ref == 0.0)                       # parity zeros are exact zeros (pyx:1324-1327)
    _check_eri(E[tuple(g["eri_cart_idx"].T)], g["eri_cart_val"])      # the reference's own values
    assert abs(E.sum() - float(g["eri_cart_sum"])) < 1e-9 * abs(float(g["eri_cart_sum"]))
    # 8-fold permutational symmetry is exact by construction
    assert np.array_equal(E, E.transpose(1, 0, 2, 3)) and np.array_equal(E, E.transpose(0, 1, 3, 2)) and np.array_equal(E, E.transpose(2, 3, 0, 1))
    c = ctx.counts()
    assert (c["unique_quartets"], c["surviving_quartets"]) == oracle.parity_surviving_quartets(g["lmn"])


@pytest.mark.parametrize("name", ["n2_ccpvtz", "ne2_uhf_ccpvqz", "et100"])
def test_engine_fill_matches_per_quartet_fill_and_is_reproducible(tb, name, monkeypatch):
    """The dense tensor from the shell-quartet engine's fill mode (the default for contracted bases, forced here) against the
    per-AO-quartet kernel (TUNA_B200_FILL_ENGINE=0): same values within the ERI tolerance, the same exact zeros, and two engine fills are
    bitwise identical (fixed-order scatter pass)."""
    g = load_golden(name)
    ctx = context_for(g)
    monkeypatch.setenv("TUNA_B200_FILL_ENGINE", "1")
    ctx.eri_fill_cart()
    E1 = ctx.eri_download(0)
    ctx.eri_fill_cart()
    E2 = ctx.eri_download(0)
    assert np.array_equal(E1, E2)
    assert np.array_equal(E1, E1.transpose(1, 0, 2, 3)) and np.array_equal(E1, E1.transpose(2, 3, 0, 1))
    monkeypatch.setenv("TUNA_B200_FILL_ENGINE", "0")
    ctx.eri_fill_cart()
    E0 = ctx.eri_download(0)
    _check_eri(E1, E0)
    assert np.array_equal(E1 == 0.0, E0 == 0.0)


def test_eri_ne2_ccpvqz_vs_compiled_reference(tb, oracle):
    """Full Cartesian tensor (140^4) against the UNMODIFIED reference engine from oracle/_ref (g functions, T up to ~3e6)."""
    g = load_golden("ne2_uhf_ccpvqz")
    ctx = context_for(g)
    ctx.eri_fill_cart()
    E = ctx.eri_download(0)
    _check_eri(E[tuple(g["eri_cart_idx"].T)], g["eri_cart_val"])
    eng = oracle.reference_engine()
    fb = oracle_basis(oracle, g)
    n = fb.ncart
    if eng is not None:
        ref = np.asarray(eng.calculate_electron_repulsion_integrals(n, np.empty((n,) * 4), oracle.reference_basis_objects(fb), oracle.max_threads()))
    else:
        ref = oracle.eri_fill(fb)
    _check_eri(E, ref)
    assert abs(np.abs(E).max() - float(g["eri_cart_max"])) < 1e-12


@pytest.mark.parametrize("name", ["n2_ccpvtz", "ne2_uhf_ccpvqz", "et100"])
def test_cart_to_sph_and_stored_jk(tb, oracle, name):
    g = load_golden(name)
    ctx = context_for(g)
    ctx.set_transform(g["U"])
    ctx.eri_fill_cart()
    ctx.eri_cart_to_sph()
    nbf = int(g["nbf"])
    Es = ctx.eri_download(1)
    assert Es.shape == (nbf,) * 4
    ref = g["eri_sph_val"]
    assert np.abs(Es[tuple(g["eri_sph_idx"].T)] - ref).max() < 2e-12
    assert abs(np.linalg.norm(Es) - float(g["eri_sph_fro"])) < 1e-11 * float(g["eri_sph_fro"])
    P = tb.workloads.fixed_density(nbf)
    J, K = ctx.jk_stored(P)
    assert np.abs(J - g["Jfix"]).max() < 1e-11 * max(1.0, np.abs(g["Jfix"]).max())
    assert np.abs(K - g["Kfix"]).max() < 1e-11 * max(1.0, np.abs(g["Kfix"]).max())
    # against the oracle's einsums on the SAME tensor: isolates the contraction kernel
    assert np.abs(J - oracle.coulomb(P, Es)).max() < 1e-11 and np.abs(K - oracle.exchange(P, Es)).max() < 1e-11
    # a general (non-symmetric) density and a stack of densities
    rng = np.random.default_rng(7)
    Ps = rng.standard_normal((3, nbf, nbf))
    Js, Ks = ctx.jk_stored(Ps)
    for d in range(3):
        assert np.abs(Js[d] - oracle.coulomb(Ps[d], Es)).max() < 1e-10
        assert np.abs(Ks[d] - oracle.exchange(Ps[d], Es)).max() < 1e-10
    # linearity
    J2, K2 = ctx.jk_stored(2.0 * Ps[0] - 0.5 * Ps[1])
    assert np.abs(J2 - (2.0 * Js[0] - 0.5 * Js[1])).max() < 1e-10 and np.abs(K2 - (2.0 * Ks[0] - 0.5 * Ks[1])).max() < 1e-10


@pytest.mark.parametrize("name", ["h2_631g", "n2_ccpvtz", "co_b3lyp_ccpvtz", "ne2_uhf_ccpvqz", "et100", "n2_ccpvtz_cartharm"])
def test_scf_sequence_stored_and_direct(tb, name):
    """The reference's own SCF sequence: every recorded P_i must give the recorded J_i, K_i (1e-11), in both modes,
    and the final J/K must reproduce the reference's Coulomb/exchange energies (tuna_scf.py:376,380 / :454-459) to 1e-10."""
    g = load_golden(name)
    ctx = context_for(g)
    ctx.set_transform(g["U"])
    ctx.eri_fill_cart()
    ctx.eri_cart_to_sph()
    keys = sorted(k for k in g.files if k.startswith("seq") and k.endswith("_P"))
    assert keys
    for k in keys:
        P, ref = g[k], g[k[:-2] + "_out"]
        kind = k.split("_")[2]
        J, K = ctx.jk_stored(P)
        got = J if kind == "J" else K
        assert np.abs(got - ref).max() < 1e-11, (k, "stored")
        Jd, Kd = ctx.jk_direct(P)
        gotd = Jd if kind == "J" else Kd
        assert np.abs(gotd - ref).max() < 1e-11, (k, "direct")
    # What a J/K error does to the energy: the reference forms E_J = 1/2 sum P_new J(P_old), E_K = -1/4 sum P_new K(P_old)
    # (tuna_scf.py:376,380 with the old-J/new-P convention of :1141; UHF analogues :454-459).  With the converged P_new on
    # both sides, our J/K on the last recorded P_old must give the same energies as the reference's recorded J/K to 1e-10.
    last = int(g["kept_iterations"][-1])
    P_new = g["P_final"]
    for k in [k for k in keys if k.startswith(f"seq{last}_")]:
        P, ref = g[k], g[k[:-2] + "_out"]
        kind = k.split("_")[2]
        J, K = ctx.jk_direct(P)
        got = J if kind == "J" else K
        assert abs(0.5 * np.sum(P_new * (got - ref))) < 1e-10, (k, "energy")


@pytest.mark.parametrize("name", ["n2_ccpvtz", "et100"])
def test_direct_jk_fixed_density_and_screening(tb, name):
    g = load_golden(name)
    ctx = context_for(g)
    ctx.set_transform(g["U"])
    P = tb.workloads.fixed_density(int(g["nbf"]))
    J0, K0 = ctx.jk_direct(P, tau=0.0)
    assert ctx.counts()["evaluated_last_direct"] == ctx.counts()["surviving_quartets"]
    scale = max(1.0, np.abs(g["Kfix"]).max())
    assert np.abs(J0 - g["Jfix"]).max() < 1e-11 * scale and np.abs(K0 - g["Kfix"]).max() < 1e-11 * scale
    J1, K1 = ctx.jk_direct(P)      # default Schwarz threshold
    assert np.abs(J1 - g["Jfix"]).max() < 1e-11 * scale and np.abs(K1 - g["Kfix"]).max() < 1e-11 * scale
    assert ctx.counts()["evaluated_last_direct"] <= ctx.counts()["surviving_quartets"]
    # Schwarz bound |(ij|kl)| <= Q_ij Q_kl on a sample
    Q = ctx.schwarz()
    idx = g["eri_cart_idx"]
    bound = Q[idx[:, 0], idx[:, 1]] * Q[idx[:, 2], idx[:, 3]]
    assert np.all(np.abs(g["eri_cart_val"]) <= bound * (1 + 1e-12) + 1e-15)
    # two-rank sharding of the quartet list: partial J/K sum to the whole
    parts = []
    for r in range(2):
        ctx.set_shard(r, 2)
        parts.append(ctx.jk_direct(P, tau=0.0))
    ctx.set_shard(0, 1)
    assert np.abs(parts[0][0] + parts[1][0] - J0).max() < 1e-11 * scale
    assert np.abs(parts[0][1] + parts[1][1] - K0).max() < 1e-11 * scale


def test_provider_reference_signatures(tb, oracle):
    """The six reference entry points, called the way tuna_kernel / tuna_scf call them."""
    g = load_golden("n2_ccpvtz")
    bfs = basis_objects(g)
    n, nbf = int(g["ncart"]), int(g["nbf"])
    # tuna_integral.calculate_electron_repulsion_integrals(n_basis, ERI_AO, bfs, num_threads): fills and returns the buffer
    buf = np.empty((n,) * 4)
    ret = tb.calculate_electron_repulsion_integrals(n, buf, bfs, 4)
    assert ret is buf
    _check_eri(buf[tuple(g["eri_cart_idx"].T)], g["eri_cart_val"])
    # single quartet, incl. an x-parity zero
    for q in [(0, 0, 0, 0), (7, 40, 1, 41), (17, 60, 18, 58), (30, 69, 22, 50)]:
        v = tb.calculate_electron_repulsion_integral(*[bfs[i] for i in q])
        assert abs(v - buf[q]) < 1e-12
    # tuna_kernel flow: two-electron integrals -> spherical transformation -> tuna_scf J/K
    calc = SimpleNamespace(cartesian_harmonics=False, number_of_threads=4, method=SimpleNamespace(method_base="HF"))
    mol = SimpleNamespace(spherical_harmonic_transformation_matrix=g["U"])
    for mode in ("stored", "direct"):
        tb.configure(mode=mode)
        h = tb.calculate_two_electron_integrals(n, bfs, calc)
        assert h.shape == (n,) * 4
        one = np.eye(n)
        S, T, V, D, Q, eri = tb.transform_to_spherical_harmonics(one, one, one, np.stack([one] * 3), np.stack([one] * 3), h, mol, calc, True)
        assert S.shape == (nbf, nbf) and D.shape == (3, nbf, nbf) and eri.shape == (nbf,) * 4
        np.testing.assert_allclose(S, g["U"] @ g["U"].T, atol=1e-14)
        P = tb.workloads.fixed_density(nbf)
        J = tb.calculate_coulomb_matrix(P, eri)
        K = tb.calculate_exchange_matrix(P, eri)
        assert np.abs(J - g["Jfix"]).max() < 1e-11 * 14 and np.abs(K - g["Kfix"]).max() < 1e-11 * 14
        assert J.flags.c_contiguous and J.dtype == np.float64
    tb.configure(mode="auto")
    # post-HF consumers need ndarray behaviour from the handle (tuna_mp.py:632 does 2*ERI - ERI.swapaxes(1,3))
    dense = np.asarray(eri)
    assert isinstance(dense, np.ndarray) and dense.shape == (nbf,) * 4
    assert np.abs(dense[tuple(g["eri_sph_idx"].T)] - g["eri_sph_val"]).max() < 2e-12
    mix = 2 * eri - eri.swapaxes(1, 3)
    assert np.allclose(mix, 2 * dense - dense.swapaxes(1, 3))
    # a plain ndarray handed to the J/K functions (what post-HF code and the unpatched glue do)
    J = tb.calculate_coulomb_matrix(P, dense)
    assert np.abs(J - g["Jfix"]).max() < 1e-11 * 14


def test_error_behaviour(tb):
    g = load_golden("h2_631g")
    bfs = basis_objects(g)
    bfs[1].origin = np.array([0.1, 0.0, 0.0])
    with pytest.raises(tb.TunaError):
        tb.calculate_electron_repulsion_integral(*bfs)
    ctx = context_for(g)
    with pytest.raises(tb.TunaError):
        ctx.jk_stored(np.eye(4))              # no tensor resident
    ctx.set_transform(g["U"])
    with pytest.raises(tb.TunaError):
        ctx.jk_direct(np.eye(5))              # wrong shape
    with pytest.raises(tb.TunaError):
        ctx.eri_single(0, 0, 0, 9)


@pytest.mark.parametrize("name", ["n2_ccpvtz", "ne2_uhf_ccpvqz"])
def test_direct_engines_agree_and_general_density(tb, oracle, name, monkeypatch):
    """Shell-quartet engine vs the per-component kernel vs the stored path, incl. a NON-symmetric density
    (K[P] = K[S] + K[A]; the reference's guess densities are asymmetric at the 1e-8 level)."""
    g = load_golden(name)
    nbf = int(g["nbf"])
    rng = np.random.default_rng(11)
    P = rng.standard_normal((2, nbf, nbf))
    P[1] = (P[1] + P[1].T) / 2
    ctx = context_for(g)
    ctx.set_transform(g["U"])
    ctx.eri_fill_cart()
    ctx.eri_cart_to_sph()
    Js, Ks = ctx.jk_stored(P)
    Jd, Kd = ctx.jk_direct(P, tau=0.0)
    monkeypatch.setenv("TUNA_B200_DIRECT_ENGINE", "generic")
    ctx2 = context_for(g)
    ctx2.set_transform(g["U"])
    Jg, Kg = ctx2.jk_direct(P, tau=0.0)
    scale = max(1.0, np.abs(Ks).max())
    for a, b in ((Jd, Js), (Kd, Ks), (Jg, Js), (Kg, Ks)):
        assert np.abs(a - b).max() < 1e-11 * scale


@pytest.mark.parametrize("name", ["ne2_uhf_ccpvqz", "et100"])
def test_direct_build_is_bitwise_reproducible(tb, name):
    """North star part 4: J/K of the direct mode do not depend on the order in which the class jobs of six streams and their CTAs arrive.
    The generation-4 engine accumulates every contribution as two 64-bit integers (fixed_split, csrc/shell_jk.cuh): integer addition is
    associative, so repeated builds - on the same context and on a fresh one - are bit-identical, and they agree with FP64-atomic
    accumulation to round-off."""
    g = load_golden(name)
    nbf = int(g["nbf"])
    rng = np.random.default_rng(21)
    P = rng.standard_normal((2, nbf, nbf))
    P = (P + P.transpose(0, 2, 1)) / 2
    ctx = context_for(g)
    ctx.set_transform(g["U"])
    J0, K0 = ctx.jk_direct(P, 1e-16)
    for _ in range(3):
        J1, K1 = ctx.jk_direct(P, 1e-16)
        assert np.array_equal(J0, J1) and np.array_equal(K0, K1)
    ctx2 = context_for(g)
    ctx2.set_transform(g["U"])
    J2, K2 = ctx2.jk_direct(P, 1e-16)
    assert np.array_equal(J0, J2) and np.array_equal(K0, K2)
    ctx2.set_shard(1, 3)                   # partial builds of a sharded run are reproducible as well
    Ja, Ka = ctx2.jk_direct(P, 1e-16)
    Jb, Kb = ctx2.jk_direct(P, 1e-16)
    assert np.array_equal(Ja, Jb) and np.array_equal(Ka, Kb)


def test_high_angular_momentum_h_shells(tb, oracle):
    """A synthetic two-centre basis with every shell type up to H (L = 5, the reference's maximum, tuna_molecule.py:612-618):
    full Cartesian tensor vs the oracle, and the shell-quartet engine (multi-chunk (hh|hh) class tables) vs the stored path."""
    from tuna_b200 import workloads as w
    from tuna_b200.basis import flatten, from_arrays
    shells_a = [(0, [1.3], [1.0]), (1, [0.9], [1.0]), (5, [1.1], [1.0])]
    shells_b = [(2, [0.8], [1.0]), (4, [1.2], [1.0]), (5, [0.7], [1.0]), (3, [1.0, 0.4], [0.6, 0.5])]
    b = w.shells_to_components([shells_a, shells_b], [0.0, 1.9])
    bfs = from_arrays(b["origins"], b["lmn"], b["nprim"], b["exps"], b["raw_coefs"])
    n = len(bfs)
    fb = oracle.FlatBasis.from_reference_objects(bfs)
    ref = oracle.eri_fill(fb)
    ctx = tb.Context(0)
    ctx.set_basis(*flatten(bfs))
    ctx.set_transform(np.eye(n))
    ctx.eri_fill_cart()
    E = ctx.eri_download(0)
    _check_eri(E, ref)
    ctx.eri_cart_to_sph()
    rng = np.random.default_rng(3)
    P = rng.standard_normal((2, n, n))
    P = (P + P.transpose(0, 2, 1)) / 2
    Js, Ks = ctx.jk_stored(P)
    Jd, Kd = ctx.jk_direct(P, tau=0.0)
    scale = max(1.0, np.abs(Ks).max())
    assert np.abs(Jd - Js).max() < 1e-11 * scale and np.abs(Kd - Ks).max() < 1e-11 * scale
    assert np.abs(Js[0] - oracle.coulomb(P[0], ref)).max() < 1e-11 * scale
    assert ctx.counts()["evaluated_last_direct"] == ctx.counts()["surviving_quartets"]


def test_permuted_and_incomplete_shell_lists(tb, oracle):
    """The provider receives a flat list of per-component Basis objects.  (i) Components of a shell need not be contiguous
    (DECONTRACT emits component-major order, tuna_molecule.py:557-565): a random permutation of the list must give the
    permuted J/K.  (ii) A list that is NOT a union of complete shells (one d component removed) cannot use the shell engine
    and must fall back to the per-component direct kernel — same answers as the stored path."""
    g = load_golden("n2_ccpvtz")
    bfs = basis_objects(g)
    n = len(bfs)
    rng = np.random.default_rng(21)
    P = rng.standard_normal((n, n))
    P = (P + P.T) / 2
    ctx = tb.Context(0)
    from tuna_b200.basis import flatten
    ctx.set_basis(*flatten(bfs))
    ctx.set_transform(np.eye(n))
    J0, K0 = ctx.jk_direct(P, tau=0.0)
    perm = rng.permutation(n)
    ctx2 = tb.Context(0)
    ctx2.set_basis(*flatten([bfs[i] for i in perm]))
    ctx2.set_transform(np.eye(n))
    J1, K1 = ctx2.jk_direct(P[np.ix_(perm, perm)], tau=0.0)
    assert np.abs(J1 - J0[np.ix_(perm, perm)]).max() < 1e-11 and np.abs(K1 - K0[np.ix_(perm, perm)]).max() < 1e-11
    # (ii) drop the first d_xx component
    drop = next(i for i, b in enumerate(bfs) if tuple(int(x) for x in b.shell) == (2, 0, 0))
    keep = [i for i in range(n) if i != drop]
    ctx3 = tb.Context(0)
    ctx3.set_basis(*flatten([bfs[i] for i in keep]))
    ctx3.set_transform(np.eye(n - 1))
    Pk = P[np.ix_(keep, keep)]
    J2, K2 = ctx3.jk_direct(Pk, tau=0.0)
    ctx3.eri_fill_cart()
    ctx3.eri_cart_to_sph()
    Js, Ks = ctx3.jk_stored(Pk)                      # n - 1 = 69 is odd: also covers the non-TMA stored kernel
    assert np.abs(J2 - Js).max() < 1e-11 and np.abs(K2 - Ks).max() < 1e-11
    E = ctx3.eri_download(1)
    assert np.abs(Js - oracle.coulomb(Pk, E)).max() < 1e-11


def test_many_densities(tb, oracle):
    """More densities than one pass handles (stored: 4 per launch; direct: limited by the shared-memory J/K blocks)."""
    g = load_golden("et100")
    ctx = context_for(g)
    ctx.set_transform(g["U"])
    ctx.eri_fill_cart()
    ctx.eri_cart_to_sph()
    nbf = int(g["nbf"])
    rng = np.random.default_rng(8)
    P = rng.standard_normal((6, nbf, nbf))
    P = (P + P.transpose(0, 2, 1)) / 2
    Js, Ks = ctx.jk_stored(P)
    Jd, Kd = ctx.jk_direct(P, tau=0.0)
    E = ctx.eri_download(1)
    for d in (0, 3, 5):
        assert np.abs(Js[d] - oracle.coulomb(P[d], E)).max() < 1e-10 and np.abs(Ks[d] - oracle.exchange(P[d], E)).max() < 1e-10
    scale = max(1.0, np.abs(Ks).max())
    assert np.abs(Jd - Js).max() < 1e-11 * scale and np.abs(Kd - Ks).max() < 1e-11 * scale


@pytest.mark.parametrize("n", [16, 22, 60, 62, 100])
def test_stored_streaming_kernels_on_uploaded_tensors(tb, n, monkeypatch):
    """Stored J/K on uploaded tensors vs the reference's two einsums (tuna_scf.py:42,70): a pair-symmetric tensor takes the
    symmetric streaming kernel (register-resident J, MT = 4 / 8 / 16 variants), an arbitrary tensor must fall back to the
    general kernels and still be exact; several (also non-symmetric) densities at once."""
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n * n, n * n))
    for E in (((A + A.T) / 2).reshape(n, n, n, n), A.reshape(n, n, n, n)):
        for nD in ((1, 3) if n <= 62 else (1,)):
            P = rng.standard_normal((nD, n, n))
            Jr = np.einsum("ijkl,dkl->dij", E, P, optimize=True)
            Kr = np.einsum("ilkj,dkl->dij", E, P, optimize=True)
            for kern in ("", "tma"):
                if kern:
                    monkeypatch.setenv("TUNA_B200_STORED_KERNEL", kern)
                else:
                    monkeypatch.delenv("TUNA_B200_STORED_KERNEL", raising=False)
                ctx = tb.Context(0)
                ctx.eri_upload(np.ascontiguousarray(E))
                J, K = ctx.jk_stored(P)
                tol = 1e-12 * n * n
                assert np.abs(np.asarray(J).reshape(Jr.shape) - Jr).max() < tol and np.abs(np.asarray(K).reshape(Kr.shape) - Kr).max() < tol
