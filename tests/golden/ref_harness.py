"""Drive the UNMODIFIED reference (h-brough/TUNA at /root/reference) programmatically.

TEST INFRASTRUCTURE, usable only in the development container: /root/reference does not exist on the
GPU box, so nothing under `-m gpu`, `smoke()` or `bench.py` imports this module.  It is what
`make_golden.py` uses to harvest the committed fixtures, and what the `not gpu` tests use (when the
reference tree is present) to cross-check the oracle restatements against the real thing.

Recipe follows SURVEY.md section 8(c): stub `termcolor`/`matplotlib`, put TUNA/ on sys.path (the
reference is a flat-module program, TUNA/tuna.py:3), and expose the compiled engine from
oracle/_ref/ as the package `tuna_integrals` the reference imports (TUNA/tuna_kernel.py:2).
"""
import importlib
import os
import sys
import time
import types
from unittest import mock

import numpy as np

REFERENCE_ROOT = os.environ.get("TUNA_REFERENCE", "/root/reference")
REPO_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF_SO_DIR = os.path.join(REPO_ROOT, "oracle", "_ref")
STAGED_TUNA_DIR = os.path.join(REF_SO_DIR, "TUNA")     # oracle/build_ref.sh stages the unmodified modules here for the GPU box

BOHR_PER_ANGSTROM = 1.8897261259065457


def reference_tuna_dir():
    """The reference's module directory: the read-only tree in the development container, else the copy staged by oracle/build_ref.sh."""
    live = os.path.join(REFERENCE_ROOT, "TUNA")
    if os.path.isdir(live):
        return live
    if os.path.isfile(os.path.join(STAGED_TUNA_DIR, "tuna_energy.py")):
        return STAGED_TUNA_DIR
    return None


def reference_available() -> bool:
    return reference_tuna_dir() is not None and any(f.startswith("tuna_integral") and f.endswith(".so") for f in os.listdir(REF_SO_DIR))


def load_reference_engine():
    """Import the compiled, unmodified reference integral engine from oracle/_ref (travels to the GPU box)."""
    if REF_SO_DIR not in sys.path:
        sys.path.insert(0, REF_SO_DIR)
    return importlib.import_module("tuna_integral")


_loaded = {}


def load_reference():
    """Import the reference's Python modules; returns a namespace with the modules the hot path touches."""
    if _loaded:
        return _loaded["ns"]
    if not reference_available():
        raise RuntimeError("reference modules not present (neither " + REFERENCE_ROOT + " nor " + STAGED_TUNA_DIR + ")")
    termcolor = types.ModuleType("termcolor")
    termcolor.colored = lambda s, *a, **k: s
    sys.modules.setdefault("termcolor", termcolor)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.font_manager",
                 "matplotlib.ticker", "matplotlib.lines", "matplotlib.patches", "matplotlib.cm"):
        sys.modules.setdefault(name, mock.MagicMock())
    engine = load_reference_engine()
    pkg = types.ModuleType("tuna_integrals")
    pkg.tuna_integral = engine
    pkg.__path__ = []
    sys.modules["tuna_integrals"] = pkg
    sys.modules["tuna_integrals.tuna_integral"] = engine
    tuna_dir = reference_tuna_dir()
    if tuna_dir not in sys.path:
        sys.path.insert(0, tuna_dir)
    ns = types.SimpleNamespace()
    ns.ints = engine
    for short, mod in (("util", "tuna_util"), ("calc", "tuna_calc"), ("energ", "tuna_energy"), ("kern", "tuna_kernel"),
                       ("scf", "tuna_scf"), ("molecule", "tuna_molecule"), ("basis", "tuna_basis")):
        setattr(ns, short, importlib.import_module(mod))
    _loaded["ns"] = ns
    return ns


def parse_line(ns, line: str):
    """Minimal restatement of the CLI split (TUNA/tuna.py:59-161): 'SPE : N N 1.10 : HF CC-PVTZ : KEYWORDS'."""
    sections = [s.strip() for s in line.split(":")]
    calc_type = sections[0].upper()
    geom = sections[1].split()
    method_string, basis = sections[2].split()
    params = sections[3].split() if len(sections) >= 4 else []
    symbols = [g.upper() for g in geom[:2] if not _is_float(g)]
    lengths = [0.0] + [float(g) for g in geom if _is_float(g)]
    coords_1d = np.array(lengths) * BOHR_PER_ANGSTROM
    coordinates = ns.util.one_dimension_to_three(coords_1d)
    method_string = method_string.upper()
    unrestricted = method_string.startswith("U")
    base = method_string[1:] if unrestricted else method_string
    method = next(m for m in ns.util.electronic_structure_methods if m.name == base)
    method.unrestricted = unrestricted
    calculation = ns.calc.Calculation(calc_type, method, time.perf_counter(), params, basis.upper(), symbols, True)
    return calculation, symbols, coordinates


def _is_float(s):
    try:
        float(s)
        return True
    except ValueError:
        return False


class Recorder:
    """Wraps the reference's J/K functions (TUNA/tuna_scf.py:27-72) to record every call of one energy evaluation."""

    def __init__(self, ns):
        self.ns = ns
        self.calls = []          # (kind, nbf, P, out)
        self.eri_cart = {}       # ncart -> tensor
        self.eri_sph = {}        # nbf -> tensor
        self.bases = {}          # ncart -> list[Basis]
        self.U = {}
        self.E_guess = None      # E handed to the main SCF loop (guess_objects[3], tuna_scf.py:1334)

    def __enter__(self):
        ns = self.ns
        self._J, self._K = ns.scf.calculate_coulomb_matrix, ns.scf.calculate_exchange_matrix
        self._two = ns.kern.calculate_two_electron_integrals
        self._sph = ns.kern.transform_to_spherical_harmonics

        def J(P, ERI):
            out = self._J(P, ERI)
            self.calls.append(("J", P.shape[0], np.array(P), np.array(out)))
            return out

        def K(P, ERI):
            out = self._K(P, ERI)
            self.calls.append(("K", P.shape[0], np.array(P), np.array(out)))
            return out

        def two(n_basis, bfs, calculation):
            out = self._two(n_basis, bfs, calculation)
            self.eri_cart[n_basis] = out
            self.bases[n_basis] = bfs
            return out

        def sph(S, T, V, D, Q, ERI_cart, molecule, calculation, silent):
            out = self._sph(S, T, V, D, Q, ERI_cart, molecule, calculation, silent)
            self.eri_sph[out[5].shape[0]] = out[5]
            self.U[out[5].shape[0]] = np.array(molecule.spherical_harmonic_transformation_matrix)
            return out

        self._loop = ns.scf.run_self_consistent_field_cycle

        def loop(molecule, calculation, integrals, V_NN, X, guess_objects, grid_container, silent):
            self.E_guess = guess_objects[3]      # the last (= main) SCF loop wins
            return self._loop(molecule, calculation, integrals, V_NN, X, guess_objects, grid_container, silent)

        ns.scf.run_self_consistent_field_cycle = loop
        ns.scf.calculate_coulomb_matrix, ns.scf.calculate_exchange_matrix = J, K
        ns.kern.calculate_two_electron_integrals, ns.kern.transform_to_spherical_harmonics = two, sph
        return self

    def __exit__(self, *exc):
        ns = self.ns
        ns.scf.calculate_coulomb_matrix, ns.scf.calculate_exchange_matrix = self._J, self._K
        ns.scf.run_self_consistent_field_cycle = self._loop
        ns.kern.calculate_two_electron_integrals, ns.kern.transform_to_spherical_harmonics = self._two, self._sph
        return False


def run_energy(line: str, record: bool = True):
    """Run one reference energy evaluation; returns (scf_output, recorder, calculation)."""
    ns = load_reference()
    calculation, symbols, coordinates = parse_line(ns, line)
    rec = Recorder(ns)
    with rec:
        result = ns.energ.evaluate_molecular_energy(calculation, symbols, coordinates, silent=True)
    return result, rec, calculation


def basis_to_arrays(bfs):
    """Flatten a list of reference `Basis` objects (tuna_integral.pyx:78-235) into plain arrays."""
    ncart = len(bfs)
    origins = np.array([np.array(b.origin) for b in bfs], dtype=np.float64).reshape(ncart, 3)
    lmn = np.array([np.array(b.shell) for b in bfs], dtype=np.int64).reshape(ncart, 3)
    nprim = np.array([int(b.num_exps) for b in bfs], dtype=np.int64)
    exps = np.concatenate([np.array(b.exps) for b in bfs])
    coefs = np.concatenate([np.array(b.coefs) for b in bfs])   # contraction-normalised (pyx:206-210)
    norms = np.concatenate([np.array(b.norm) for b in bfs])    # primitive norms (pyx:190)
    return dict(origins=origins, lmn=lmn, nprim=nprim, exps=exps, coefs=coefs, norms=norms)
