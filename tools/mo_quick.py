"""AO->MO transformation: device time, end-to-end time with pinned / pageable result buffers (development aid, one short GPU call)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tuna_b200
from util import load_golden, context_for
out = {}
for name in ("n2_ccpvtz", "ne2_uhf_ccpvqz"):
    g = load_golden(name)
    ctx = context_for(g); ctx.set_transform(g["U"]); ctx.eri_fill_cart(); ctx.eri_cart_to_sph()
    n = int(g["nbf"])
    C = np.linalg.qr(np.random.default_rng(1).standard_normal((n, n)))[0]
    h = tuna_b200.ERIHandle(ctx, n, "sph", "stored")
    r = {"n": n}
    for pinned in ("1", "0"):
        os.environ["TUNA_B200_PINNED_RESULTS"] = pinned
        T = tuna_b200.transform_ERI_AO_to_MO(h, C, None, True)
        ts, ks = [], []
        for _ in range(3):
            t = time.perf_counter(); T = tuna_b200.transform_ERI_AO_to_MO(h, C, None, True); ts.append(time.perf_counter() - t)
            ks.append(ctx.last_kernel_ms(4))
        r["e2e_ms_pinned" + pinned] = 1e3 * min(ts); r["kernel_ms"] = float(np.median(ks))
        t = time.perf_counter(); E = ctx.eri_download(1); r["download_ms_pinned" + pinned] = 1e3 * (time.perf_counter() - t)
    r["tflops"] = 8.0 * n ** 5 / (r["kernel_ms"] * 1e-3) / 1e12
    r["checksum"] = float(T.sum())
    out[name] = r
    print(json.dumps(out), flush=True)
