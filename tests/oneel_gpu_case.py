"""One GPU case of the one-electron path (run by tests/test_zz_fullsize.py::test_one_electron_integrals_gpu in a child process):
the CUDA launch, the C ABI and the reference's Python signatures against the oracle and the reference's own fixtures."""
import os
import sys
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def main(name):
    import tuna_b200
    from oracle import tuna_oracle as oracle
    from test_one_electron import NAMES, close, molecule_inputs, sub_basis
    from util import basis_objects, load_golden, oracle_basis
    g = load_golden(name)
    fb = oracle_basis(oracle, g)
    zs, ch, origin = molecule_inputs(fb, name)
    bfs = basis_objects(g)
    atoms = [SimpleNamespace(origin=np.array([0.0, 0.0, z]), charge=float(c)) for z, c in zip(zs, ch)]
    got = tuna_b200.calculate_one_electron_integrals(len(bfs), bfs, len(atoms), atoms, origin, 4)
    ref = oracle.one_electron(fb, zs, ch, origin)
    assert len(got) == 5 and got[3].shape == (3, fb.ncart, fb.ncart)
    for nm, a, b in zip(NAMES, got, ref):
        close(a, b, f"{name} {nm}")
    U = np.array(g["U"])                          # the reference's own spherical matrices of this configuration
    assert np.abs(U @ got[0] @ U.T - np.array(g["S"])).max() < 1e-12
    assert np.abs(U @ got[1] @ U.T - np.array(g["T"])).max() < 1e-11 and np.abs(U @ got[2] @ U.T - np.array(g["V_NE"])).max() < 1e-11
    idx = np.arange(0, fb.ncart, 3)
    sub = sub_basis(oracle, fb, idx)
    S12 = tuna_b200.calculate_cross_basis_overlap_matrix(len(bfs), len(idx), bfs, [bfs[i] for i in idx], 4)
    close(S12, oracle.cross_overlap(fb, sub), f"{name} S_cross")
    try:
        tuna_b200.calculate_one_electron_integrals(len(bfs), bfs, 1, [SimpleNamespace(origin=np.array([0.1, 0.0, 0.0]), charge=1.0)], origin, 4)
    except tuna_b200.TunaError:
        pass
    else:
        raise AssertionError("off-axis atom was accepted")
    print("ok", name)


if __name__ == "__main__":
    main(sys.argv[1])
