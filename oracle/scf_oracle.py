"""Restated Hartree-Fock SCF loop of the reference — TEST INFRASTRUCTURE ONLY (never imported by tuna_b200/).

Purpose: the north star asks for "converged SCF total energy within 1e-10 Eh with identical iteration counts".  The
reference's Python cannot travel to the GPU box, so its SCF loop is restated here and PINNED on the CPU
(tests/test_scf_closed_loop.py): driven by the oracle's J/K it must reproduce the reference's recorded energy and
iteration count for every fixture.  On the GPU box the same loop is driven by the tuna_b200 provider.

Restated from TUNA/tuna_scf.py of h-brough/TUNA v0.12.0 (HF only: no XC, no external fields):
    density / diagonalisation       :183-250        SCF changes + convergence test   :261-335
    RHF / UHF energy                :344-410, :415-490
    RHF / UHF Fock matrices         :497-531, :542-589
    Zerner-Hehenberger damping      :763-868        DIIS error + extrapolation       :879-1061
    RHF / UHF cycle                 :1072-1162, :1165-1289     driver loop            :1292-1435
The J/K provider is injected: jk(P) -> (J, K) stands for scf.calculate_coulomb_matrix / calculate_exchange_matrix.
"""
import numpy as np


def symmetrise(M):                                   # tuna_util.py:748-764
    return 0.5 * (M + M.T)


def density_matrix(C, n_occ, n_per_orbital):         # tuna_scf.py:183-211
    occ = C[:, :n_occ]
    return symmetrise(n_per_orbital * occ @ occ.T)


def diagonalise(F, X):                               # tuna_scf.py:222-250
    eps, vec = np.linalg.eigh(symmetrise(X.T @ F @ X))
    return eps, X @ vec


def _mulliken(P, S, partition_ranges):               # tuna_scf.py:790-812
    PS = P @ S
    pops, start = [], 0
    for n in partition_ranges[:2]:
        pops.append(sum(PS[i, i] for i in range(start, start + n)))
        start += n
    while len(pops) < 2:
        pops.append(0)
    return np.array(pops)


def apply_damping(P_new, P_old_damped, commutator, cfg, P_old_before, P_very_old_damped, S, step):   # tuna_scf.py:763-868
    factor = 0
    pr = cfg["partition_ranges"]
    if cfg["damping"]:
        if cfg["damping_factor"] is not None:
            factor = float(cfg["damping_factor"])
        elif commutator > 0.01 and step > 1:
            a_out = _mulliken(P_new, S, pr)
            a1_in = _mulliken(P_old_damped, S, pr)
            a1_out = _mulliken(P_old_before, S, pr)
            a2_in = _mulliken(P_very_old_damped, S, pr)
            den = a_out - a1_out - a1_in + a2_in
            alpha = (a_out - a1_out) / den if den.all() != 0 else [0, 0]
            if len(pr) == 2:
                factor = (alpha[0] * pr[0] + alpha[1] * pr[1]) / (pr[0] + pr[1])
            else:
                factor = alpha[0] * pr[0]
            factor = max(factor, 0)
            factor = factor if factor < min(cfg["max_damping"], 1) else cfg["max_damping"]
    return factor * P_old_damped + (1 - factor) * P_new, factor


def diis_error(Fa, Fb, Pa, Pb, S, X, errors, focks, cfg):                  # tuna_scf.py:879-948
    def comm(F, P):
        e = X.T @ (F @ P @ S - S @ P @ F) @ X
        return np.mean(e * e) ** 0.5, e
    ca, ea = comm(Fa, Pa)
    cb, eb = comm(Fb, Pb)
    errors.append(np.concatenate((ea.flatten(), eb.flatten())))
    focks.append((Fa, Fb))
    if len(focks) > cfg["max_DIIS_matrices"]:
        del focks[0]
        del errors[0]
    return max(ca, cb), ca, cb


def apply_diis(commutator, step, P, Pa, Pb, focks, errors, na, nb, X, n_per_orbital, cfg):   # tuna_scf.py:960-1061
    if step > 2 and cfg["DIIS"] and commutator < 0.3:
        n = len(errors)
        E = np.array(errors)
        B = np.empty((n + 1, n + 1))
        B[:n, :n] = E @ E.T
        B[:n, -1] = -1
        B[-1, :n] = -1
        B[-1, -1] = 0
        rhs = np.zeros(n + 1)
        rhs[-1] = -1
        Pa_d = Pb_d = None
        try:
            c = np.linalg.solve(B, rhs)[:n]
            Fa = np.tensordot(c, np.array([f[0] for f in focks]), axes=(0, 0))
            Fb = np.tensordot(c, np.array([f[1] for f in focks]), axes=(0, 0))
            Pa_d = density_matrix(diagonalise(Fa, X)[1], na, n_per_orbital)
            Pb_d = density_matrix(diagonalise(Fb, X)[1], nb, n_per_orbital)
        except np.linalg.LinAlgError:
            focks.clear()
            errors.clear()
        if Pa_d is not None and Pb_d is not None:
            Pa, Pb = symmetrise(Pa_d), symmetrise(Pb_d)
            P = symmetrise(Pa + Pb) / 2
    return P, Pa, Pb


def run_scf(jk, fx, max_iter=100):
    """fx: fixture mapping with S, T, V_NE, X, the guess densities, occupations and the reference's settings.
    Returns (total energy, iterations, final P)."""
    S, T, V, X = (np.array(fx[k]) for k in ("S", "T", "V_NE", "X"))
    cfg = dict(damping=bool(fx["damping"]), damping_factor=None if np.isnan(float(fx["damping_factor"])) else float(fx["damping_factor"]),
               max_damping=float(fx["max_damping"]), DIIS=bool(fx["DIIS"]), max_DIIS_matrices=int(fx["max_DIIS_matrices"]),
               partition_ranges=[int(x) for x in fx["partition_ranges"]])
    conv = {k: float(fx["conv_" + k]) for k in ("delta_E", "max_DP", "RMS_DP", "commutator")}
    hfx, V_NN = float(fx["HFX_prop"]), float(fx["V_NN"])
    unrestricted = bool(fx["unrestricted"])
    n_alpha, n_beta, n_docc = int(fx["n_alpha"]), int(fx["n_beta"]), int(fx["n_doubly_occ"])
    P, Pa, Pb, E = np.array(fx["P_guess"]), np.array(fx["P_guess_alpha"]), np.array(fx["P_guess_beta"]), float(fx["E_guess"])
    z = np.zeros_like(P)
    P_old, Pa_old, Pb_old, P_bd, Pa_bd, Pb_bd = z, z, z, z, z, z
    errors, focks = [], []
    for step in range(1, max_iter + 1):
        E_old = E
        if not unrestricted:                                                   # run_restricted_SCF_cycle, :1072-1162
            P_very_old, P_old_bd, P_old = P_old, P_bd, P
            J, K = jk(P)
            F = symmetrise(T + V + J - 0.5 * K * hfx)
            commutator, _, _ = diis_error(F, F, P, P, S, X, errors, focks, cfg)
            P = density_matrix(diagonalise(F, X)[1], n_docc, 2)
            E = np.einsum("ij,ij->", P, T) + np.einsum("ij,ij->", P, V) + 0.5 * np.einsum("ij,ij->", P, J) - 0.25 * np.einsum("ij,ij->", P, K) * hfx
            P, _, _ = apply_diis(commutator, step, P, P / 2, P / 2, focks, errors, n_docc, n_docc, X, 2, cfg)
            P_bd = P
            P, _ = apply_damping(P, P_old, commutator, cfg, P_old_bd, P_very_old, S, step)
        else:                                                                  # run_unrestricted_SCF_cycle, :1165-1289
            Pa_very_old, Pb_very_old, Pa_old_bd, Pb_old_bd = Pa_old, Pb_old, Pa_bd, Pb_bd
            P_old, Pa_old, Pb_old = P, Pa, Pb
            Ja, Ka = jk(Pa)
            Jb, Kb = jk(Pb)
            Fa = symmetrise(T + V + Ja + Jb - Ka * hfx)
            Fb = symmetrise(T + V + Ja + Jb - Kb * hfx)
            commutator, ca, cb = diis_error(Fa, Fb, Pa, Pb, S, X, errors, focks, cfg)
            Pa = density_matrix(diagonalise(Fa, X)[1], n_alpha, 1)
            Pb = density_matrix(diagonalise(Fb, X)[1], n_beta, 1)
            Pt = Pa + Pb
            E = (np.einsum("ij,ij->", Pt, T) + np.einsum("ij,ij->", Pt, V) + 0.5 * np.einsum("ij,ij->", Pt, Ja + Jb)
                 - 0.5 * np.einsum("ij,ij->", Pa, Ka) * hfx - 0.5 * np.einsum("ij,ij->", Pb, Kb) * hfx)
            _, Pa, Pb = apply_diis(commutator, step, P, Pa, Pb, focks, errors, n_alpha, n_beta, X, 1, cfg)
            Pa_bd, Pb_bd = Pa, Pb
            Pa, _ = apply_damping(Pa, Pa_old, ca, cfg, Pa_old_bd, Pa_very_old, S, step)
            Pb, _ = apply_damping(Pb, Pb_old, cb, cfg, Pb_old_bd, Pb_very_old, S, step)
            P = Pa + Pb
        dP = P - P_old                                                         # calculate_SCF_changes, :261-288
        dE, maxDP, rmsDP = E - E_old, np.max(np.abs(dP)), np.mean(dP ** 2) ** 0.5
        if abs(dE) < conv["delta_E"] and abs(maxDP) < conv["max_DP"] and abs(rmsDP) < conv["RMS_DP"] and abs(commutator) < conv["commutator"]:
            return E + V_NN, step, P
    raise RuntimeError("SCF not converged")        # the reference errors out as well (tuna_scf.py:1435)
