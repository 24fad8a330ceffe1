"""Sweep the group-size heuristics of the shell-quartet engine (development aid)."""
import os, sys, time, json, subprocess
import numpy as np
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import tuna_b200
    from tuna_b200 import workloads as w
    from tuna_b200.basis import flatten, from_arrays
    nbf = int(sys.argv[2])
    b = w.even_tempered_diatomic(nbf)
    bfs = from_arrays(b["origins"], b["lmn"], b["nprim"], b["exps"], b["raw_coefs"])
    ctx = tuna_b200.Context(0)
    ctx.set_basis(*flatten(bfs)); ctx.set_transform(np.eye(len(bfs)))
    P = w.fixed_density(len(bfs))
    for _ in range(2):
        ctx.jk_direct(P, 1e-16)
    print(json.dumps({"ms": ctx.last_kernel_ms(3), "launches": ctx.counts()["launches"]}))
else:
    for nbf in (200, 400):
        for gdiv in (8,):
            for spl in (128, 256):
                env = dict(os.environ, TUNA_B200_G_DIV=str(gdiv), TUNA_B200_SMEM_PER_LANE=str(spl))
                r = subprocess.run([sys.executable, __file__, "child", str(nbf)], env=env, capture_output=True, text=True)
                print(nbf, gdiv, spl, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
