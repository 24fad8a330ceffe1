"""The kernel math (tuna_b200/csrc/eri_core.cuh + pairtable.hpp, `__host__ __device__`) compiled for the CPU and
checked against the oracle.  This is a development aid for a container without a GPU — the package never loads it."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from util import load_golden, oracle_basis

HERE = os.path.dirname(os.path.abspath(__file__))


VARIANTS = {"": []}      # the shipped kernel bodies (no compile-time variants are left)


class LazyEmul:
    """Compiles tests/host_emul/emul.cpp with the variant's macros on first use (skipped variant cases cost nothing)."""

    def __init__(self, tag):
        self.variant = tag
        self._lib = None

    def __getattr__(self, name):
        if self._lib is None:
            tag = self.variant
            so = os.path.join(HERE, "host_emul", f"libemul{'_' + tag if tag else ''}.so")
            src = os.path.join(HERE, "host_emul", "emul.cpp")
            subprocess.run(["g++", "-O2", "-fopenmp", "-fPIC", "-shared", "-x", "c++"] + VARIANTS[tag] + ["-o", so, src, "-lm"], check=True)
            lib = ctypes.CDLL(so)
            lib.emul_boys.restype = ctypes.c_double
            lib.emul_boys.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double]
            self._lib = lib
        return getattr(self._lib, name)


@pytest.fixture(scope="module", params=list(VARIANTS))
def emul(request):
    return LazyEmul(request.param)


def test_boys_kernel_function(emul, oracle):
    rng = np.random.default_rng(3)
    Ts = np.concatenate([[0.0, 1e-12, 0.03125, 39.999, 40.0, 40.001, 1e3, 3e6, 1e13], rng.uniform(0, 45, 200), 10 ** rng.uniform(-6, 6, 100)])
    worst = 0.0
    for T in Ts:
        for M in (0, 1, 5, 12, 20):
            for m in {0, M // 2, M}:
                got, ref = emul.emul_boys(M, m, float(T)), oracle.boys(m, float(T))
                worst = max(worst, abs(got - ref) / abs(ref))
    assert worst < 2e-14, worst
    assert emul.emul_boys(20, 20, 0.0) == 1.0 / 41.0 and emul.emul_boys(20, 0, 0.0) == 1.0      # T == 0 branch is exact (pyx:1557-1563)


@pytest.mark.parametrize("name", ["h2_631g", "n2_ccpvtz", "et100"])
def test_quartet_math_vs_oracle(emul, oracle, name):
    g = load_golden(name)
    fb = oracle_basis(oracle, g)
    n = fb.ncart
    dp, ip, lp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int64)
    oz = np.ascontiguousarray(fb.origins[:, 2])
    lmn = np.ascontiguousarray(fb.lmn, dtype=np.int32)
    npr = np.ascontiguousarray(fb.nprim, dtype=np.int32)
    off = np.ascontiguousarray(fb.offsets, dtype=np.int64)
    ceff = np.ascontiguousarray(fb.coefs * fb.norms)
    out = np.empty((n,) * 4)
    emul.emul_eri_fill(n, oz.ctypes.data_as(dp), lmn.ctypes.data_as(ip), npr.ctypes.data_as(ip), off.ctypes.data_as(lp),
                       fb.exps.ctypes.data_as(dp), ceff.ctypes.data_as(dp), out.ctypes.data_as(dp))
    ref = oracle.eri_fill(fb)
    assert np.all(np.abs(out - ref) <= np.maximum(1e-12, 1e-13 * np.abs(ref)))
    assert np.array_equal(out == 0.0, ref == 0.0)


@pytest.mark.parametrize("name,nb,budget,target", [("h2_631g", 2, None, 16), ("n2_ccpvtz", 2, None, 16), ("n2_ccpvtz", 1, 96, 4), ("n2_ccpvtz", 4, 300, 0)])
def test_engine_fill_mode_vs_oracle(emul, oracle, name, nb, budget, target, monkeypatch):
    """Dense Cartesian tensor through the engine's fill mode (scratch rows per work item + the fixed-order scatter pass) against
    the oracle: contracted shells split into primitive chunks, several integral-buffer chunks per class, 1 / 2 / 4 quartets per batch."""
    monkeypatch.setenv("TUNA_EMUL_NB", str(nb))
    monkeypatch.setenv("TUNA_EMUL_PSPLIT_TARGET", str(target))
    if budget:
        monkeypatch.setenv("TUNA_EMUL_IT_BUDGET", str(budget))
    g = load_golden(name)
    fb = oracle_basis(oracle, g)
    n = fb.ncart
    dp, ip, lp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int64)
    oz = np.ascontiguousarray(fb.origins[:, 2])
    lmn = np.ascontiguousarray(fb.lmn, dtype=np.int32)
    npr = np.ascontiguousarray(fb.nprim, dtype=np.int32)
    off = np.ascontiguousarray(fb.offsets, dtype=np.int64)
    ceff = np.ascontiguousarray(fb.coefs * fb.norms)
    out = np.full((n,) * 4, np.nan)
    rc = emul.emul_fill_shell4(n, oz.ctypes.data_as(dp), lmn.ctypes.data_as(ip), npr.ctypes.data_as(ip), off.ctypes.data_as(lp),
                               fb.exps.ctypes.data_as(dp), ceff.ctypes.data_as(dp), out.ctypes.data_as(dp))
    assert rc == 0
    ref = oracle.eri_fill(fb)
    assert np.all(np.abs(out - ref) <= np.maximum(1e-12, 1e-13 * np.abs(ref)))
    assert np.array_equal(out == 0.0, ref == 0.0)
    assert np.array_equal(out, out.transpose(1, 0, 2, 3)) and np.array_equal(out, out.transpose(2, 3, 0, 1))      # one value for all eight images


ENGINES = {"gen4": "emul_jk_shell4"}


def engine_entry(emul, gen):
    """The serial CPU build of the engine body (shell4.cuh: same source as the sm_100a kernel, HostPolicy instead of DevPolicy)."""
    return getattr(emul, ENGINES[gen])


@pytest.mark.parametrize("gen", list(ENGINES))
@pytest.mark.parametrize("name", ["h2_631g", "n2_ccpvtz", "et100"])
def test_shell_engine_vs_oracle(emul, oracle, name, gen):
    """shell_jk.cuh (the direct-mode engine) run serially on the CPU: J/K for two symmetric densities vs the oracle's
    einsums over the oracle's Cartesian tensor, with and without Schwarz screening."""
    g = load_golden(name)
    fb = oracle_basis(oracle, g)
    n = fb.ncart
    E = oracle.eri_fill(fb)
    dp, ip, lp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int64)
    oz = np.ascontiguousarray(fb.origins[:, 2])
    lmn = np.ascontiguousarray(fb.lmn, dtype=np.int32)
    npr = np.ascontiguousarray(fb.nprim, dtype=np.int32)
    off = np.ascontiguousarray(fb.offsets, dtype=np.int64)
    ceff = np.ascontiguousarray(fb.coefs * fb.norms)
    rng = np.random.default_rng(1)
    P = rng.standard_normal((2, n, n))
    P = (P + P.transpose(0, 2, 1)) / 2
    J, K, stats = np.zeros_like(P), np.zeros_like(P), np.zeros(8, dtype=np.int64)
    for tau in (0.0, 1e-16):
        rc = engine_entry(emul, gen)(n, oz.ctypes.data_as(dp), lmn.ctypes.data_as(ip), npr.ctypes.data_as(ip), off.ctypes.data_as(lp),
                                fb.exps.ctypes.data_as(dp), ceff.ctypes.data_as(dp), 2, P.ctypes.data_as(dp), J.ctypes.data_as(dp),
                                K.ctypes.data_as(dp), ctypes.c_double(tau), stats.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)))
        assert rc == 0
        for d in range(2):
            Jr, Kr = oracle.coulomb(P[d], E), oracle.exchange(P[d], E)
            assert np.abs(J[d] - Jr).max() < 1e-11 and np.abs(K[d] - Kr).max() < 1e-11


@pytest.mark.parametrize("target,ket", [("0", "1"), ("5", "1"), ("16", "1"), ("16", "0"), ("7", "1")])
def test_shell_engine_primitive_split(emul, oracle, target, ket, monkeypatch):
    """Contracted classes: a shell quartet split into work items of bra primitive pairs - and, once every bra pair is its own item, of ket
    primitive pairs (ket = 1) - each digesting and flushing its partial integrals, gives the same J/K as the unsplit walk (target 0).
    N2/cc-pVTZ has shells of up to 8 primitives, i.e. 4096 primitive quartets in (ss|ss); target 7 leaves ragged last chunks on both sides."""
    monkeypatch.setenv("TUNA_EMUL_PSPLIT_TARGET", target)
    monkeypatch.setenv("TUNA_EMUL_KSPLIT", ket)
    g = load_golden("n2_ccpvtz")
    fb = oracle_basis(oracle, g)
    n = fb.ncart
    E = oracle.eri_fill(fb)
    dp, ip, lp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int64)
    oz = np.ascontiguousarray(fb.origins[:, 2])
    lmn = np.ascontiguousarray(fb.lmn, dtype=np.int32)
    npr = np.ascontiguousarray(fb.nprim, dtype=np.int32)
    off = np.ascontiguousarray(fb.offsets, dtype=np.int64)
    ceff = np.ascontiguousarray(fb.coefs * fb.norms)
    P = np.random.default_rng(4).standard_normal((1, n, n))
    P = (P + P.transpose(0, 2, 1)) / 2
    J, K, stats = np.zeros_like(P), np.zeros_like(P), np.zeros(8, dtype=np.int64)
    rc = emul.emul_jk_shell4(n, oz.ctypes.data_as(dp), lmn.ctypes.data_as(ip), npr.ctypes.data_as(ip), off.ctypes.data_as(lp), fb.exps.ctypes.data_as(dp),
                             ceff.ctypes.data_as(dp), 1, P.ctypes.data_as(dp), J.ctypes.data_as(dp), K.ctypes.data_as(dp), ctypes.c_double(0.0),
                             stats.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)))
    assert rc == 0
    assert np.abs(J[0] - oracle.coulomb(P[0], E)).max() < 1e-11 and np.abs(K[0] - oracle.exchange(P[0], E)).max() < 1e-11


def test_primitive_split_covers_every_primitive_quartet_once(emul):
    """shell4_split + the header decode of shell4_unit: the work items of a shell quartet visit every (bra, ket) primitive pair exactly once,
    an item holds at most ~target primitive quartets once bra and ket are both split, and target 0 never splits."""
    out = (ctypes.c_int * 4)()
    for nab in (1, 2, 3, 5, 9, 16, 24, 64, 81):
        for ncd in (1, 2, 3, 9, 17, 24, 64, 81):
            for target in (0, 1, 4, 7, 16, 100):
                for ket in (0, 1):
                    emul.emul_split(nab, ncd, target, ket, out)
                    psplit, clen, ksplit, klen = out[0], out[1], out[2], out[3]
                    assert psplit % ksplit == 0 and psplit >= 1 and clen >= 1 and klen >= 1
                    if target == 0:
                        assert (psplit, clen, ksplit, klen) == (1, nab, 1, ncd)
                    if not ket:
                        assert ksplit == 1 and klen == ncd
                    seen = np.zeros((nab, ncd), dtype=int)
                    largest = 0
                    for pchunk in range(psplit):
                        ia0, ic0 = (pchunk // ksplit) * clen, (pchunk % ksplit) * klen      # as packed into Quartet4::ia0
                        assert ia0 < 65536 and ic0 < 32768
                        cnt = 0
                        for it in range(clen):
                            for kk in range(klen):
                                if ia0 + it < nab and ic0 + kk < ncd:
                                    seen[ia0 + it, ic0 + kk] += 1
                                    cnt += 1
                        largest = max(largest, cnt)
                    assert (seen == 1).all(), (nab, ncd, target, ket)
                    if target > 0 and ket:
                        assert largest <= 2 * target, (nab, ncd, target, largest)      # ceil effects only


@pytest.mark.parametrize("gen,nb,budget", [("gen4", None, None), ("gen4", "1", "1500"), ("gen4", "4", "700")])
def test_shell_engine_h_shells(emul, oracle, gen, nb, budget, monkeypatch):
    """All shell types up to H, including the multi-chunk (hh|hh) class tables, on a synthetic two-centre basis; generation 4 also with
    one and four quartets per batch and with small chunk budgets (several chunks per class, table reload)."""
    if nb:
        monkeypatch.setenv("TUNA_EMUL_NB", nb)
        monkeypatch.setenv("TUNA_EMUL_IT_BUDGET", budget)
    from tuna_b200 import workloads as w
    from tuna_b200.basis import from_arrays
    shells_a = [(0, [1.3], [1.0]), (1, [0.9], [1.0]), (5, [1.1], [1.0])]
    shells_b = [(2, [0.8], [1.0]), (4, [1.2], [1.0]), (5, [0.7], [1.0]), (3, [1.0, 0.4], [0.6, 0.5])]
    b = w.shells_to_components([shells_a, shells_b], [0.0, 1.9])
    fb = oracle.FlatBasis.from_reference_objects(from_arrays(b["origins"], b["lmn"], b["nprim"], b["exps"], b["raw_coefs"]))
    n = fb.ncart
    E = oracle.eri_fill(fb)
    dp, ip, lp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int64)
    oz = np.ascontiguousarray(fb.origins[:, 2])
    lmn = np.ascontiguousarray(fb.lmn, dtype=np.int32)
    npr = np.ascontiguousarray(fb.nprim, dtype=np.int32)
    off = np.ascontiguousarray(fb.offsets, dtype=np.int64)
    ceff = np.ascontiguousarray(fb.coefs * fb.norms)
    P = np.random.default_rng(1).standard_normal((1, n, n))
    P = (P + P.transpose(0, 2, 1)) / 2
    J, K, stats = np.zeros_like(P), np.zeros_like(P), np.zeros(8, dtype=np.int64)
    rc = engine_entry(emul, gen)(n, oz.ctypes.data_as(dp), lmn.ctypes.data_as(ip), npr.ctypes.data_as(ip), off.ctypes.data_as(lp),
                                 fb.exps.ctypes.data_as(dp), ceff.ctypes.data_as(dp), 1, P.ctypes.data_as(dp), J.ctypes.data_as(dp),
                                 K.ctypes.data_as(dp), ctypes.c_double(0.0), stats.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)))
    assert rc == 0 and stats[0] == 7
    assert np.abs(J[0] - oracle.coulomb(P[0], E)).max() < 1e-11 and np.abs(K[0] - oracle.exchange(P[0], E)).max() < 1e-11


@pytest.mark.parametrize("nb,budget", [(None, None), ("1", "1500")])
def test_engine_fill_mode_h_shells(emul, oracle, nb, budget, monkeypatch):
    """Fill mode on the synthetic two-centre basis with every shell type up to H (one contracted f shell): the multi-chunk (hh|hh)-type
    classes write their scratch rows chunk by chunk; with the default budgets and with small ones (more chunks per class)."""
    if nb:
        monkeypatch.setenv("TUNA_EMUL_NB", nb)
        monkeypatch.setenv("TUNA_EMUL_IT_BUDGET", budget)
    from tuna_b200 import workloads as w
    from tuna_b200.basis import from_arrays
    shells_a = [(0, [1.3], [1.0]), (1, [0.9], [1.0]), (5, [1.1], [1.0])]
    shells_b = [(2, [0.8], [1.0]), (4, [1.2], [1.0]), (5, [0.7], [1.0]), (3, [1.0, 0.4], [0.6, 0.5])]
    b = w.shells_to_components([shells_a, shells_b], [0.0, 1.9])
    fb = oracle.FlatBasis.from_reference_objects(from_arrays(b["origins"], b["lmn"], b["nprim"], b["exps"], b["raw_coefs"]))
    n = fb.ncart
    ref = oracle.eri_fill(fb)
    dp, ip, lp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int64)
    oz = np.ascontiguousarray(fb.origins[:, 2])
    lmn = np.ascontiguousarray(fb.lmn, dtype=np.int32)
    npr = np.ascontiguousarray(fb.nprim, dtype=np.int32)
    off = np.ascontiguousarray(fb.offsets, dtype=np.int64)
    ceff = np.ascontiguousarray(fb.coefs * fb.norms)
    out = np.full((n,) * 4, np.nan)
    rc = emul.emul_fill_shell4(n, oz.ctypes.data_as(dp), lmn.ctypes.data_as(ip), npr.ctypes.data_as(ip), off.ctypes.data_as(lp),
                               fb.exps.ctypes.data_as(dp), ceff.ctypes.data_as(dp), out.ctypes.data_as(dp))
    assert rc == 0
    assert np.all(np.abs(out - ref) <= np.maximum(1e-12, 1e-13 * np.abs(ref)))
    assert np.array_equal(out == 0.0, ref == 0.0)
    assert np.array_equal(out, out.transpose(1, 0, 2, 3)) and np.array_equal(out, out.transpose(2, 3, 0, 1))


@pytest.mark.parametrize("gen", list(ENGINES))
def test_shell_engine_extreme_shells_of_the_headline_basis(emul, oracle, gen):
    """The tightest and the most diffuse shell of every angular momentum of the ET800 set (s exponent 2.1e5 ... h exponent 1.0) on both
    atoms, unit-pair densities: sampled J/K elements against single integrals of the oracle (no dense tensor needed).  The same check
    runs on the GPU at the full nbf 400 / 800 sizes (tests/test_zz_fullsize.py)."""
    from tuna_b200 import workloads as w
    from tuna_b200.basis import from_arrays
    from util import check_unit_pair_jk, pick_function, unit_pair_density
    sh = w.even_tempered_shells(800)
    sel = []
    for L in range(6):
        ex = [a for (l, a) in sh if l == L]
        sel += [(L, [a], [1.0]) for a in sorted({ex[0], ex[-1]})]
    b = w.shells_to_components([sel, sel], [0.0, 1.10 * w.BOHR_PER_ANGSTROM])
    fb = oracle.FlatBasis.from_reference_objects(from_arrays(b["origins"], b["lmn"], b["nprim"], b["exps"], b["raw_coefs"]))
    n = fb.ncart
    pairs = [(pick_function(fb, 0, 5, True, 0), pick_function(fb, 0, 5, True, 3)), (pick_function(fb, 0, 3, True, 9), pick_function(fb, 1, 5, False, 20))]
    P = np.stack([unit_pair_density(n, k, l) for k, l in pairs])
    dp, ip, lp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int64)
    oz = np.ascontiguousarray(fb.origins[:, 2])
    lmn = np.ascontiguousarray(fb.lmn, dtype=np.int32)
    npr = np.ascontiguousarray(fb.nprim, dtype=np.int32)
    off = np.ascontiguousarray(fb.offsets, dtype=np.int64)
    ceff = np.ascontiguousarray(fb.coefs * fb.norms)
    J, K, stats = np.zeros_like(P), np.zeros_like(P), np.zeros(8, dtype=np.int64)
    rc = engine_entry(emul, gen)(n, oz.ctypes.data_as(dp), lmn.ctypes.data_as(ip), npr.ctypes.data_as(ip), off.ctypes.data_as(lp),
                                 fb.exps.ctypes.data_as(dp), ceff.ctypes.data_as(dp), 2, P.ctypes.data_as(dp), J.ctypes.data_as(dp),
                                 K.ctypes.data_as(dp), ctypes.c_double(1e-16), stats.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)))
    assert rc == 0
    for d, (k, l) in enumerate(pairs):
        assert check_unit_pair_jk(oracle, fb, k, l, J[d], K[d], n_samples=200) < 1e-12
