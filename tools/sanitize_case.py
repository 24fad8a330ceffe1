"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): every kernel family once."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import tuna_b200
from util import load_golden, context_for
for name in sys.argv[1:] or ["h2_631g", "n2_ccpvtz"]:
    g = load_golden(name)
    ctx = context_for(g); ctx.set_transform(g["U"])
    ctx.eri_fill_cart(); ctx.eri_cart_to_sph()
    n = int(g["nbf"])
    P = tuna_b200.workloads.fixed_density(n)
    J, K = ctx.jk_stored(np.stack([P, P.T @ P]))
    Jd, Kd = ctx.jk_direct(P, 1e-16)
    print(name, float(np.abs(Jd - J[0]).max()), float(np.abs(Kd - K[0]).max()), ctx.eri_single(0, 0, 0, 0))
    ctx.close()
