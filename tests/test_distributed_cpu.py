"""world_size-2 test of the multi-GPU path's host logic on CPU (gloo): the round-robin chunk sharding of the unique
quartet list (tuna_b200.distributed.shard_chunks, mirrored by the CUDA kernels) and the one-all-reduce-per-build
plumbing.  Each rank folds only its share of the oracle's integrals into partial J/K with the same 8-fold-symmetric
update rule the kernels use; the all-reduced result must equal the dense einsums of the reference (tuna_scf.py:27-72)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _partial_jk(E, P, rank, world, chunk):
    """Unique quartets (i>=j, k>=l, ij>=kl) dealt to ranks in chunks; digestion as in k_jk_direct / shell_quartet."""
    from tuna_b200.distributed import shard_chunks
    n = E.shape[0]
    pairs = [(i, j) for i in range(n) for j in range(i + 1)]
    quartets = [(a, b) for a in range(len(pairs)) for b in range(a + 1)]
    Jacc, Kacc = np.zeros((n, n)), np.zeros((n, n))
    mine = 0
    for c in shard_chunks(len(quartets), rank, world, chunk):
        for a, b in quartets[c * chunk:(c + 1) * chunk]:
            (i, j), (k, l) = pairs[a], pairs[b]
            v = E[i, j, k, l] * (0.5 if i == j else 1.0) * (0.5 if k == l else 1.0) * (0.5 if a == b else 1.0)
            Jacc[i, j] += v * (P[k, l] + P[l, k]); Jacc[k, l] += v * (P[i, j] + P[j, i])
            Kacc[i, l] += v * P[k, j]; Kacc[j, l] += v * P[k, i]; Kacc[i, k] += v * P[l, j]; Kacc[j, k] += v * P[l, i]
            mine += 1
    return Jacc + Jacc.T, Kacc + Kacc.T, mine


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import tuna_oracle as orc
    from util import load_golden, oracle_basis
    g = load_golden("h2_631g")
    E = orc.eri_fill(oracle_basis(orc, g), 1)
    P = np.array(g["P_final"])
    J, K, mine = _partial_jk(E, P, rank, world, chunk=8)
    t = torch.from_numpy(np.stack([J, K]))
    dist.all_reduce(t)                               # the one collective of a Fock build
    cnt = torch.tensor([mine])
    dist.all_reduce(cnt)
    if rank == 0:
        np.save(out, np.concatenate([t.numpy().ravel(), [float(cnt.item())], orc.coulomb(P, E).ravel(), orc.exchange(P, E).ravel()]))
    dist.destroy_process_group()


def test_two_rank_sharded_fock_build(tmp_path):
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(2, 29533, out), nprocs=2, join=True)
    r = np.load(out)
    J, K = r[:16].reshape(4, 4), r[16:32].reshape(4, 4)
    assert int(r[32]) == 55                              # every unique quartet of H2/6-31G evaluated exactly once
    Jr, Kr = r[33:49].reshape(4, 4), r[49:65].reshape(4, 4)
    assert np.abs(J - Jr).max() < 1e-13 and np.abs(K - Kr).max() < 1e-13


def test_shard_chunks_partition():
    from tuna_b200.distributed import shard_chunks
    for n in (0, 1, 63, 64, 65, 1000, 4097):
        for world in (1, 2, 4, 8):
            seen = sorted(c for r in range(world) for c in shard_chunks(n, r, world))
            assert seen == list(range((n + 127) // 128))
