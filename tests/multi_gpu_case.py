"""Worker of tests/test_multi_gpu.py: one process per GPU (torch.distributed.run, NCCL), the sharded direct Fock build of
tuna_b200.distributed.FockBuilder checked on rank 0 against the single-GPU build, the reference's recorded J/K and (Ne2) the
reference's converged energy / iteration count.  Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import tuna_b200
    from tuna_b200 import workloads as w
    from tuna_b200.basis import flatten, from_arrays
    from tuna_b200.distributed import FockBuilder
    from util import basis_objects, load_golden
    case = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = {"case": case, "world": world}
    if case == "ne2":
        g = load_golden("ne2_uhf_ccpvqz")
        bfs, U = basis_objects(g), np.array(g["U"])
        P2 = np.stack([np.array(g["P_alpha_final"]), np.array(g["P_beta_final"])])
        Pfix = w.fixed_density(int(g["nbf"]))
    else:
        nbf = int(case[2:])
        b = w.even_tempered_diatomic(nbf)
        bfs = from_arrays(b["origins"], b["lmn"], b["nprim"], b["exps"], b["raw_coefs"])
        U = np.eye(len(bfs))
        Pfix = w.fixed_density(len(bfs))
        P2 = np.stack([Pfix, Pfix.T @ Pfix / len(bfs)])
    ctx = tuna_b200.Context(local)
    ctx.set_basis(*flatten(bfs))
    ctx.set_transform(U)
    fb1 = FockBuilder(ctx, nD=1, tau=1e-16)
    J1, K1 = fb1.build(Pfix)
    fb2 = FockBuilder(ctx, nD=2, tau=1e-16)
    J2, K2 = fb2.build(P2)
    J2b, K2b = fb2.build(P2)                      # a second, identical build (run-to-run reproducibility of the sharded path)
    if rank == 0:
        solo = tuna_b200.Context(local)           # the same work on ONE GPU, unsharded
        solo.set_basis(*flatten(bfs))
        solo.set_transform(U)
        Js1, Ks1 = solo.jk_direct(Pfix, 1e-16)
        Js2, Ks2 = solo.jk_direct(P2, 1e-16)
        scale = max(1.0, float(np.abs(Ks2).max()))
        res.update(dJ1=float(np.abs(J1 - Js1).max()), dK1=float(np.abs(K1 - Ks1).max()), dJ2=float(np.abs(J2 - Js2).max()), dK2=float(np.abs(K2 - Ks2).max()),
                   rerun=float(max(np.abs(J2 - J2b).max(), np.abs(K2 - K2b).max())), scale=scale)
        if case == "ne2":
            res.update(dJfix=float(np.abs(J1 - g["Jfix"]).max()), dKfix=float(np.abs(K1 - g["Kfix"]).max()))
        solo.close()
    if case == "ne2":
        # the restated SCF loop (oracle/scf_oracle.py, pinned to the reference's energies and iteration counts on the CPU) runs on every
        # rank; each of its Fock builds (J/K of the alpha, then of the beta density, tuna_scf.py:571-577) is sharded over all ranks
        from oracle import scf_oracle
        energy, iterations, _ = scf_oracle.run_scf(lambda P: fb1.build(P), g)
        res.update(dE=float(energy - float(g["energy"])), iterations=int(iterations), ref_iterations=int(g["n_iterations"]))
    dist.barrier()
    if rank == 0:
        print("RESULT " + json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
