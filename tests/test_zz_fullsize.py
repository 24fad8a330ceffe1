"""Direct ERI + J/K at BASELINE.json's FULL sizes (configs[4]: even-tempered N2, nbf 400 and 800 — the bench's headline workload).

The reference cannot hold 8 ncart^4 bytes at these sizes, so parity is checked without a dense tensor:
  * unit-pair densities P = e_k e_l^T + e_l e_k^T turn J and K into single integrals, J_ij = (ij|kl) + (ij|lk),
    K_ij = (il|kj) + (ik|lj), which the oracle evaluates one by one (pyx:1376-1414 restated in oracle/eri_oracle.c); sampled (i, j)
    cover every angular class up to (hh|hh) and exponents from 0.1 to 2.1e5.  Tolerance: SURVEY.md 8(d)'s sweep tolerance
    max(1e-12, 1e-13 |ERI|) per integral;
  * size-independent properties on the fixed synthetic density: J and K symmetric, linear in P, additive over the two shards of
    a 2-GPU run (tuna_set_shard), and the quartet accounting (evaluated <= parity-surviving <= unique).
The same unit-pair check runs on the CPU for the extreme shells of this basis through the host emulation
(tests/test_host_emul.py::test_shell_engine_extreme_shells_of_the_headline_basis).  File name: runs last under `-x`."""
import numpy as np
import pytest

from util import check_unit_pair_jk, pick_function, unit_pair_density

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nbf", [400, 800])
def test_direct_jk_full_size(oracle, nbf):
    import tuna_b200
    from tuna_b200 import workloads as w
    from tuna_b200.basis import flatten, from_arrays
    b = w.even_tempered_diatomic(nbf)
    bfs = from_arrays(b["origins"], b["lmn"], b["nprim"], b["exps"], b["raw_coefs"])
    fb = oracle.FlatBasis.from_reference_objects(bfs)
    n = fb.ncart
    assert n == {400: 524, 800: 1102}[nbf]
    ctx = tuna_b200.Context(0)
    ctx.set_basis(*flatten(bfs))
    ctx.set_transform(np.eye(n))                      # Cartesian components (CARTHARM, tuna_kernel.py:481-483): J/K element = AO integral sums
    tau = 1e-16

    # ---- parity against the oracle, element by element ----
    pairs = [(pick_function(fb, 0, 5, True, 0), pick_function(fb, 0, 5, True, 3)),        # tight h, same centre
             (pick_function(fb, 0, 3, True, 9), pick_function(fb, 1, 5, False, 20)),      # tight f on A, diffuse h on B
             (pick_function(fb, 0, 0, False), pick_function(fb, 1, 4, True, 14))]         # diffuse s on A, tight g on B
    P = np.stack([unit_pair_density(n, k, l) for k, l in pairs])
    J, K = ctx.jk_direct(P, tau)
    for d, (k, l) in enumerate(pairs):
        check_unit_pair_jk(oracle, fb, k, l, J[d], K[d], n_samples=250, seed=11 + d)

    # ---- size-independent properties on the bench's density ----
    c = ctx.counts()
    unique, surviving = oracle.parity_surviving_quartets(fb.lmn)
    assert (c["unique_quartets"], c["surviving_quartets"]) == (unique, surviving)
    P1 = w.fixed_density(n)
    A = np.random.default_rng(7).standard_normal((n, n))
    P2 = (A + A.T) / 2
    (J1, J2), (K1, K2) = ctx.jk_direct(np.stack([P1, P2]), tau)
    evaluated = ctx.counts()["evaluated_last_direct"]
    assert 0 < evaluated <= surviving
    sj, sk = np.abs(J1).max(), np.abs(K1).max()
    assert np.abs(J1 - J1.T).max() <= 1e-11 * sj and np.abs(K1 - K1.T).max() <= 1e-11 * sk
    Jc, Kc = ctx.jk_direct(P1 + 2.0 * P2, tau)
    assert np.abs(Jc - (J1 + 2.0 * J2)).max() <= 1e-11 * max(sj, np.abs(Jc).max())
    assert np.abs(Kc - (K1 + 2.0 * K2)).max() <= 1e-11 * max(sk, np.abs(Kc).max())
    parts = []
    for r in range(2):
        ctx.set_shard(r, 2)
        parts.append(ctx.jk_direct(P1, tau))
    ctx.set_shard(0, 1)
    assert np.abs(parts[0][0] + parts[1][0] - J1).max() <= 1e-11 * sj
    assert np.abs(parts[0][1] + parts[1][1] - K1).max() <= 1e-11 * sk
    ctx.close()


# ---------------------------------------------------------------------------------------------------------------------
# One-electron integrals on the GPU.  Written after the round's GPU budget was spent: the kernel math is
# parity-checked on the CPU (tests/test_one_electron.py); these cases exercise the CUDA launch, the C ABI and the reference's
# Python signatures.  Tolerance as in tests/test_one_electron.py.
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["h2_631g", "n2_ccpvtz", "ne2_uhf_ccpvqz", "et100"])
def test_one_electron_integrals_gpu(name):
    """Runs tests/oneel_gpu_case.py in a child process: this CUDA path has never run on a GPU, and a crash in it must not take
    the rest of the suite's report down."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, "oneel_gpu_case.py"), name], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, f"child failed (rc {r.returncode}):\n{r.stdout[-1500:]}\n{r.stderr[-3000:]}"
