"""Quick device-side timings of every kernel on a few workloads (development aid; not the bench)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import tuna_b200  # noqa: E402
from tuna_b200 import workloads as w  # noqa: E402
from tuna_b200.basis import flatten, from_arrays  # noqa: E402
from util import basis_objects, load_golden  # noqa: E402

out = {}
ctx0 = tuna_b200.Context(0)
out["fp64_peak_tflops"] = ctx0.fp64_peak_probe()
print("fp64 peak probe TFLOP/s:", out["fp64_peak_tflops"], flush=True)


def run(name, bfs, U, stored=True, direct=True):
    r = {}
    ctx = tuna_b200.Context(0)
    t = time.time(); ctx.set_basis(*flatten(bfs)); r["set_basis_s"] = time.time() - t
    ctx.set_transform(U)
    c = ctx.counts(); r.update(c)
    r["alg_eri_flops"], r["alg_digest_flops"] = ctx.algorithmic_flops()
    nbf = U.shape[0]
    P = w.fixed_density(nbf)
    if stored:
        for _ in range(2):
            ctx.eri_fill_cart(); r["eri_fill_ms"] = ctx.last_kernel_ms(0)
            t = time.time(); ctx.eri_cart_to_sph(); r["sph_ms"] = ctx.last_kernel_ms(1); r["sph_wall_s"] = time.time() - t
        for _ in range(3):
            t = time.time(); J, K = ctx.jk_stored(P); r["jk_stored_wall_ms"] = (time.time() - t) * 1e3
            r["jk_stored_ms"] = ctx.last_kernel_ms(2)
    if direct:
        for tau in (0.0, 1e-16):
            for _ in range(2):
                t = time.time(); Jd, Kd = ctx.jk_direct(P, tau); wall = time.time() - t
            r[f"jk_direct_ms_tau{tau}"] = ctx.last_kernel_ms(3)
            r[f"jk_direct_wall_ms_tau{tau}"] = wall * 1e3
            r[f"evaluated_tau{tau}"] = ctx.counts()["evaluated_last_direct"]
        if stored:
            r["direct_vs_stored_J"] = float(np.abs(Jd - J).max()); r["direct_vs_stored_K"] = float(np.abs(Kd - K).max())
    ctx.close()
    out[name] = r
    print(name, json.dumps(r), flush=True)


for name in ("n2_ccpvtz", "ne2_uhf_ccpvqz"):
    g = load_golden(name)
    run(name, basis_objects(g), g["U"])
for nbf in (100, 200, 400):
    b = w.even_tempered_diatomic(nbf)
    bfs = from_arrays(b["origins"], b["lmn"], b["nprim"], b["exps"], b["raw_coefs"])
    run(f"et{nbf}", bfs, np.eye(len(bfs)), stored=(nbf <= 200), direct=True)
json.dump(out, open("gpurun_out/calib.json", "w"), indent=1)
