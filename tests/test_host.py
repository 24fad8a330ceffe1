"""CPU tests of the host logic and the C-ABI boundary (no compute calls: there is no GPU here)."""
import os
import re

import numpy as np
import pytest

from util import basis_objects, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """libtuna_b200.so loads and exports exactly what include/tuna_b200.h declares."""
    from tuna_b200 import _lib
    from tuna_b200.build import build
    build()
    header = open(os.path.join(ROOT, "include", "tuna_b200.h")).read()
    declared = set(re.findall(r"\b(tuna_[a-z0-9_]+)\s*\(", header)) - {"tuna_ctx"}
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.EXPORTS)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import tuna_b200
    with pytest.raises(tuna_b200.TunaError):
        tuna_b200.Context(0)
    g = load_golden("h2_631g")
    with pytest.raises(tuna_b200.TunaError):
        tuna_b200.calculate_electron_repulsion_integral(*basis_objects(g))
    with pytest.raises(tuna_b200.TunaError):                      # no CPU path behind the AO->MO transformation either
        tuna_b200.transform_ERI_AO_to_MO(np.zeros((4,) * 4), np.eye(4), None, True)
    with pytest.raises(tuna_b200.TunaError):
        tuna_b200.calculate_coulomb_matrix(np.eye(4), np.zeros((4,) * 4))


def test_basis_mirror_normalisation():
    """tuna_b200.Basis reproduces Basis.normalize (pyx:174-210) on the reference's own numbers."""
    from tuna_b200.basis import Basis, flatten
    for name in ("h2_631g", "n2_ccpvtz", "ne2_uhf_ccpvqz"):
        g = load_golden(name)
        off = np.concatenate([[0], np.cumsum(g["nprim"])])
        bfs = []
        for i in range(int(g["ncart"])):
            s = slice(off[i], off[i + 1])
            b = Basis(g["origins"][i], g["lmn"][i], int(g["nprim"][i]), g["exps"][s], g["coefs"][s])
            np.testing.assert_allclose(b.norm, g["norms"][s], rtol=1e-14)
            np.testing.assert_allclose(b.coefs, g["coefs"][s], rtol=1e-13)     # normalising normalised coefficients is idempotent
            bfs.append(b)
        oz, lmn, nprim, exps, ceff = flatten(bfs)
        np.testing.assert_allclose(ceff, g["norms"] * g["coefs"], rtol=1e-13)
        assert lmn.dtype == np.int32 and oz.shape == (int(g["ncart"]),)
    raw = Basis([0, 0, 0], [0, 0, 0], 1, [1.0], [1.0])
    assert abs(raw.norm[0] - 0.71270547) < 1e-8 and abs(raw.coefs[0] - 1.0) < 1e-14


def test_flatten_rejects_off_axis_centres():
    from tuna_b200.basis import Basis, flatten
    with pytest.raises(ValueError):
        flatten([Basis([0.0, 0.2, 0.0], [0, 0, 0], 1, [1.0], [1.0])])


def test_even_tempered_workloads(oracle):
    from tuna_b200 import workloads as w
    expect = {100: (112, 381.4697265625), 200: (242, None), 400: (524, 26214.4), 800: (1102, None)}
    for nbf, (ncart, amax) in expect.items():
        b = w.even_tempered_diatomic(nbf)
        assert len(b["nprim"]) == ncart and w.spherical_count(b["lmn"]) == nbf
        if amax:
            assert abs(b["exps"].max() - amax) < 1e-9
    # the nbf=100 point is the basis of the et100 fixture (which came through the reference's CUSTOM basis reader)
    g = load_golden("et100")
    b = w.even_tempered_diatomic(100)
    np.testing.assert_array_equal(b["lmn"], g["lmn"])
    np.testing.assert_allclose(b["exps"], g["exps"], rtol=1e-9)      # the CUSTOM reader round-trips exponents through text
    np.testing.assert_allclose(b["origins"], g["origins"], rtol=1e-15)
    P = w.fixed_density(5)
    assert np.array_equal(P, P.T)
    # quartet counts of SURVEY.md 8(d): 2.00e7 unique / 5.37e6 surviving for ET100
    u, s = oracle.parity_surviving_quartets(b["lmn"])
    assert abs(u / 2.00e7 - 1) < 0.01 and abs(s / 5.37e6 - 1) < 0.01


def test_handle_forwards_ndarray_protocol():
    from tuna_b200.provider import ERIHandle
    h = ERIHandle(None, 3, "sph", "stored")
    h._host = np.arange(81.0).reshape(3, 3, 3, 3)
    assert h.shape == (3, 3, 3, 3) and h.ndim == 4
    assert np.array_equal(np.asarray(h), h._host)
    assert np.array_equal(2 * h - h.swapaxes(1, 3), 2 * h._host - h._host.swapaxes(1, 3))
    assert np.array_equal(np.einsum("ijkl,kl->ij", h, np.eye(3)), np.einsum("ijkl,kl->ij", h._host, np.eye(3)))
    assert h[1, 2, 0, 1] == h._host[1, 2, 0, 1]


def test_install_rebinds_reference_names_and_uninstall_restores():
    """install() on the UNMODIFIED reference modules (dev container only): the ten names of INTEGRATION.md section 2 are rebound,
    the reference's TunaError / timer / log are adopted, uninstall() restores the originals.  No compute call (no GPU here)."""
    import importlib
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import ref_harness as rh
    if not rh.reference_available():
        pytest.skip("reference tree not present (GPU box)")
    import tuna_b200
    from tuna_b200 import _lib, provider
    ns = rh.load_reference()
    ci = importlib.import_module("tuna_ci")
    before = {"eri": ns.ints.calculate_electron_repulsion_integrals, "J": ns.scf.calculate_coulomb_matrix, "mo": ci.transform_ERI_AO_to_MO}
    saved_err = _lib.error_class
    originals = tuna_b200.install()
    try:
        assert len(originals) == 10
        assert ns.ints.calculate_one_electron_integrals is tuna_b200.calculate_one_electron_integrals
        assert ns.ints.calculate_cross_basis_overlap_matrix is tuna_b200.calculate_cross_basis_overlap_matrix
        assert ns.ints.calculate_electron_repulsion_integrals is tuna_b200.calculate_electron_repulsion_integrals
        assert ns.ints.calculate_electron_repulsion_integral is tuna_b200.calculate_electron_repulsion_integral
        assert ns.kern.calculate_two_electron_integrals is tuna_b200.calculate_two_electron_integrals
        assert ns.kern.transform_to_spherical_harmonics is tuna_b200.transform_to_spherical_harmonics
        assert ns.scf.calculate_coulomb_matrix is tuna_b200.calculate_coulomb_matrix
        assert ns.scf.calculate_exchange_matrix is tuna_b200.calculate_exchange_matrix
        assert ci.transform_ERI_AO_to_MO is tuna_b200.transform_ERI_AO_to_MO and ci.transform_ERI_AO_to_SO is tuna_b200.transform_ERI_AO_to_SO
        assert _lib.error_class is ns.util.TunaError and provider._timer_fn is ns.util.timer and provider._log_fn is ns.util.log
        # same positional signatures as the functions they replace
        import inspect
        for (modname, name), orig in originals.items():
            try:
                ref_params = list(inspect.signature(orig).parameters)
            except (TypeError, ValueError):
                continue                                            # cpdef functions of the compiled engine expose no signature
            assert list(inspect.signature(getattr(tuna_b200, name)).parameters) == ref_params, name
    finally:
        tuna_b200.uninstall(originals)
        _lib.error_class = saved_err
    assert ns.ints.calculate_electron_repulsion_integrals is before["eri"] and ns.scf.calculate_coulomb_matrix is before["J"]
    assert ci.transform_ERI_AO_to_MO is before["mo"]


def test_mode_selection_follows_the_reference_dispatch():
    """`auto` must keep the dense tensor exactly when kern.run_post_SCF_energy_calculation will hand ERI_AO to a consumer that
    needs an ndarray (tuna_kernel.py:1146-1175): checked with the reference's own Calculation / Method objects."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import ref_harness as rh
    if not rh.reference_available():
        pytest.skip("reference tree not present (GPU box)")
    from tuna_b200 import provider
    ns = rh.load_reference()
    names = {m.name for m in ns.util.electronic_structure_methods}
    expect = {"HF": False, "B3LYP": False, "MP2": True, "CCSD": True, "CIS": True}
    checked = 0
    for method, dense in expect.items():
        if method not in names:
            continue
        calc, _, _ = rh.parse_line(ns, f"SPE : H H 0.74 : {method} 6-31G")
        assert provider._dense_needed(calc) is dense, method
        checked += 1
    assert checked >= 3
    # memory rule of `auto` (no post-HF consumer): stored while 2.2 x 8 n^4 bytes fit 40 % of the free device memory
    calc, _, _ = rh.parse_line(ns, "SPE : H H 0.74 : HF 6-31G")
    saved = provider._free_device_bytes
    try:
        provider._free_device_bytes = lambda: 10 * 2 ** 30
        provider.configure(mode="auto")
        assert provider._choose_mode(70, calc) == "stored" and provider._choose_mode(140, calc) == "direct"
        calc_mp2, _, _ = rh.parse_line(ns, "SPE : H H 0.74 : MP2 6-31G") if "MP2" in names else (calc, None, None)
        if "MP2" in names:
            assert provider._choose_mode(140, calc_mp2) == "stored"
        provider.configure(mode="direct")
        assert provider._choose_mode(4, calc) == "direct"
    finally:
        provider._free_device_bytes = saved
        provider.configure(mode="auto")


def test_engine_sass_has_no_floating_point_atomics():
    """North star part 4 / DESIGN.md section 4: the shell-quartet engine accumulates J/K as 64-bit INTEGER words (order independent, bitwise
    reproducible); its kernels must contain no FP64 atomics.  (`k_jk_direct`, the per-component fallback for bases that do not group into
    full shells, still uses them and is excluded.)  The AO->MO step must sit on the FP64 tensor cores (SASS DMMA)."""
    import shutil
    import subprocess
    from tuna_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    fn, bad, seen_engine, dmma = "", set(), False, False
    for line in sass.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
            seen_engine = seen_engine or "k_shell4_" in fn
        elif "k_shell4_" in fn and ("RED" in line or "ATOM" in line) and "F64" in line:
            bad.add(fn)
        elif "k_axis_gemm" in fn and "DMMA" in line:
            dmma = True
    assert seen_engine and not bad, sorted(bad)[:3]
    assert dmma
