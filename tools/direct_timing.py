"""Direct-mode timing of one even-tempered workload (development aid, used by the GPU-call scripts of this directory):

    python tools/direct_timing.py child <nbf>

prints one JSON line: best kernel time of four direct Fock builds (ms, CUDA events around all class-job launches) plus two weighted
checksums of J and K for a quick parity cross-check between builds / environment knobs (TUNA_B200_NB, TUNA_B200_SMEM_PER_LANE,
TUNA_B200_TERM_MAX, TUNA_B200_DBG_SKIP ...)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    import tuna_b200
    from tuna_b200 import workloads as w
    from tuna_b200.basis import flatten, from_arrays
    nbf = int(sys.argv[2] if sys.argv[1] == "child" else sys.argv[1])
    b = w.even_tempered_diatomic(nbf)
    bfs = from_arrays(b["origins"], b["lmn"], b["nprim"], b["exps"], b["raw_coefs"])
    ctx = tuna_b200.Context(0)
    ctx.set_basis(*flatten(bfs))
    ctx.set_transform(np.eye(len(bfs)))
    P = w.fixed_density(len(bfs))
    best = 1e30
    for _ in range(4):
        J, K = ctx.jk_direct(P, 1e-16)
        best = min(best, ctx.last_kernel_ms(3))
    W = np.random.default_rng(5).standard_normal(J.shape[-2:])
    print(json.dumps({"ms": best, "cj": float((np.asarray(J).reshape(W.shape) * W).sum()), "ck": float((np.asarray(K).reshape(W.shape) * W).sum()),
                      "nj": float(np.abs(J).max()), "nk": float(np.abs(K).max())}))
