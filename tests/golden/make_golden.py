"""Generate the committed golden fixtures by running the UNMODIFIED reference in this container.

    python tests/golden/make_golden.py [name ...]

Needs /root/reference and oracle/_ref (oracle/build_ref.sh).  Writes tests/golden/<name>.npz.  Each file holds
the main-basis inputs of one BASELINE.json config (flattened Basis list, U), the reference's own results for
it (ERI samples and checksums in both bases, J/K on the fixed density of SURVEY.md 8(d), the recorded SCF
sequence P_i -> J_i, K_i, final energy and iteration count) and nothing the GPU box cannot use without the
reference tree.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_harness as rh  # noqa: E402
from tuna_b200 import workloads  # noqa: E402

CONFIGS = {
    "h2_631g": "SPE : H H 0.74 : HF 6-31G",
    "h2_631g_nodiis": "SPE : H H 0.74 : HF 6-31G : NODIIS",
    "n2_ccpvtz": "SPE : N N 1.10 : HF CC-PVTZ",
    "n2_ccpvtz_cartharm": "SPE : N N 1.10 : HF CC-PVTZ : CARTHARM",
    "co_b3lyp_ccpvtz": "SPE : C O 1.128 : B3LYP CC-PVTZ",
    "ne2_uhf_ccpvqz": "SPE : NE NE 3.1 : UHF CC-PVQZ : NOROTATE",
    "et100": "SPE : N N 1.10 : HF CUSTOM : BASIS {basis_file}",
}
N_SAMPLES = 4000
KEEP_ITERATIONS = 4     # SCF iterations whose P/J/K are stored (first two, middle, last)


def make(name):
    line = CONFIGS[name]
    cwd = os.getcwd()
    tmp = None
    if "{basis_file}" in line:
        # the CLI upper-cases the whole line (TUNA/tuna.py:87): use an upper-case relative path
        tmp = tempfile.mkdtemp()
        with open(os.path.join(tmp, "ET100.TUNA"), "w") as f:
            f.write(workloads.even_tempered_basis_file(100))
        os.chdir(tmp)
        line = line.format(basis_file="ET100.TUNA")
    try:
        (out, molecule, energy, P_final), rec, calc = rh.run_energy(line)
    finally:
        os.chdir(cwd)
    ncart = molecule.n_cartesian_basis
    eri_cart = rec.eri_cart[ncart]
    bfs = rec.bases[ncart]
    flat = rh.basis_to_arrays(bfs)
    U = np.array(molecule.spherical_harmonic_transformation_matrix)
    nbf = U.shape[0]
    eri_sph = rec.eri_sph.get(nbf, eri_cart)     # CARTHARM: no rotation happened (tuna_kernel.py:481-483)
    main_calls = [c for c in rec.calls if c[1] == nbf]
    # the STO-3G guess SCF (tuna_energy.py:278-282) has a different nbf in every config, so nbf selects the main SCF
    unrestricted = bool(calc.method.unrestricted)
    per_iter = 4 if len([c for c in main_calls[:4] if c[0] == "J"]) == 2 and main_calls[1][0] == "J" else 2
    n_iter = len(main_calls) // per_iter
    keep = sorted(set([0, 1, n_iter // 2, n_iter - 1]))[:KEEP_ITERATIONS]
    seq = {}
    for it in keep:
        for c_i, call in enumerate(main_calls[it * per_iter:(it + 1) * per_iter]):
            kind, _, P, M = call
            seq[f"seq{it}_{c_i}_{kind}_P"] = P
            seq[f"seq{it}_{c_i}_{kind}_out"] = M

    # inputs of the main SCF loop (for oracle/scf_oracle.py): guess densities = the densities of the first recorded calls
    first = main_calls[:per_iter]
    if per_iter == 4:      # UHF call order: J(Pa), J(Pb), K(Pa), K(Pb)  (tuna_scf.py:571-577)
        Pg_a, Pg_b = first[0][2], first[1][2]
        Pg = Pg_a + Pg_b
    else:
        Pg = first[0][2]
        Pg_a = Pg_b = Pg / 2
    ns0 = rh.load_reference()
    V_NN = float(ns0.kern.calculate_nuclear_repulsion_energy(molecule.charges, molecule.coordinates, calc, True)) if calc.diatomic else 0.0
    scf_inputs = dict(P_guess=Pg, P_guess_alpha=Pg_a, P_guess_beta=Pg_b, E_guess=float(rec.E_guess) if rec.E_guess is not None else 0.0, V_NN=V_NN,
                      n_alpha=int(molecule.n_alpha), n_beta=int(molecule.n_beta), n_doubly_occ=int(molecule.n_doubly_occ),
                      partition_ranges=np.array(molecule.partition_ranges, dtype=np.int64), damping=bool(calc.damping),
                      damping_factor=float("nan") if calc.damping_factor is None else float(calc.damping_factor),
                      max_damping=float(calc.max_damping), DIIS=bool(calc.DIIS), max_DIIS_matrices=int(calc.max_DIIS_matrices),
                      conv_delta_E=calc.SCF_conv["delta_E"], conv_max_DP=calc.SCF_conv["max_DP"], conv_RMS_DP=calc.SCF_conv["RMS_DP"],
                      conv_commutator=calc.SCF_conv["commutator"], is_dft=bool(calc.DFT_calculation))

    rng = np.random.default_rng(12345)
    idx_c = rng.integers(0, ncart, size=(N_SAMPLES, 4))
    idx_s = rng.integers(0, nbf, size=(N_SAMPLES, 4))
    Pfix = workloads.fixed_density(nbf)
    ns = rh.load_reference()
    Jfix = ns.scf.calculate_coulomb_matrix(Pfix, eri_sph)
    Kfix = ns.scf.calculate_exchange_matrix(Pfix, eri_sph)
    # raw (un-normalised) contraction coefficients: undo Basis.normalize's common scale is impossible in general,
    # so store what the reference holds (normalised coefs + primitive norms); re-normalising them is idempotent.
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        line=np.array(CONFIGS[name]), ncart=ncart, nbf=nbf, U=U,
        origins=flat["origins"], lmn=flat["lmn"], nprim=flat["nprim"], exps=flat["exps"], coefs=flat["coefs"], norms=flat["norms"],
        eri_cart_idx=idx_c, eri_cart_val=eri_cart[tuple(idx_c.T)], eri_cart_sum=eri_cart.sum(), eri_cart_fro=np.linalg.norm(eri_cart),
        eri_cart_max=np.abs(eri_cart).max(),
        eri_sph_idx=idx_s, eri_sph_val=eri_sph[tuple(idx_s.T)], eri_sph_sum=eri_sph.sum(), eri_sph_fro=np.linalg.norm(eri_sph),
        eri_sph_max=np.abs(eri_sph).max(),
        Jfix=Jfix, Kfix=Kfix,
        energy=float(energy), n_iterations=n_iter, calls_per_iteration=per_iter, kept_iterations=np.array(keep),
        S=np.array(out.S), T=np.array(out.T), V_NE=np.array(out.V_NE), X=np.array(out.X), P_final=np.array(out.P),
        P_alpha_final=np.array(out.P_alpha), P_beta_final=np.array(out.P_beta),
        coulomb_energy=float(out.coulomb_energy), exchange_energy=float(out.exchange_energy),
        HFX_prop=float(calc.HFX_prop), unrestricted=unrestricted,
        **scf_inputs, **seq)
    print(f"{name}: ncart={ncart} nbf={nbf} E={float(energy):.12f} iterations={n_iter} calls/iter={per_iter} "
          f"sum_cart={eri_cart.sum():.9f} fro_sph={np.linalg.norm(eri_sph):.9f}")


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(CONFIGS)):
        make(n)
