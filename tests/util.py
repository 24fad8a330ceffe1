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
