"""Stored J/K kernel variants: correctness vs numpy einsum on the device tensor + timing (development aid)."""
import os, sys, json, subprocess
import numpy as np
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    import torch
    import tuna_b200
    from util import load_golden, context_for
    name = sys.argv[2]
    g = load_golden(name)
    ctx = context_for(g); ctx.set_transform(g["U"]); ctx.eri_fill_cart(); ctx.eri_cart_to_sph()
    n = int(g["nbf"])
    E = ctx.eri_download(1)
    rng = np.random.default_rng(5)
    for nD in (1, 2):
        P = rng.standard_normal((nD, n, n))
        J, K = ctx.jk_stored(P)
        Jr = np.einsum("ijkl,dkl->dij", E, P, optimize=True); Kr = np.einsum("ilkj,dkl->dij", E, P, optimize=True)
        print("nD", nD, "J err", np.abs(J - Jr).max(), "K err", np.abs(K - Kr).max(), flush=True)
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    flush2 = torch.ones(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    clean = os.environ.get("FLUSH_CLEAN", "1") == "1"
    P = rng.standard_normal((1, n, n))
    ts = []
    for _ in range(10):
        flush.add_(1.0)
        if clean:
            sink = torch.sum(flush2)
        torch.cuda.synchronize()
        ctx.jk_stored(P); ts.append(ctx.last_kernel_ms(2))
    print(json.dumps({"name": name, "n": n, "kernel": os.environ.get("TUNA_B200_STORED_KERNEL", "sym"), "ms_median": float(np.median(ts)), "ms_min": float(min(ts)),
                      "GBps": 8.0 * n ** 4 / (np.median(ts) * 1e-3) / 1e9}))
elif len(sys.argv) > 1 and sys.argv[1] == "profile":
    # ncu target (tools/gpu_round.sh): one stored J/K build and one AO->MO transformation of a golden workload (default Ne2 UHF/cc-pVQZ, 1.17 GB)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    import tuna_b200
    from util import load_golden, context_for
    wname = sys.argv[2] if len(sys.argv) > 2 else "ne2_uhf_ccpvqz"
    g = load_golden(wname)
    ctx = context_for(g); ctx.set_transform(g["U"]); ctx.eri_fill_cart(); ctx.eri_cart_to_sph()
    n = int(g["nbf"])
    rng = np.random.default_rng(5)
    P = rng.standard_normal((2 if bool(g["unrestricted"]) else 1, n, n)); P = P + P.transpose(0, 2, 1)
    J, K = ctx.jk_stored(P)
    print("stored ms", ctx.last_kernel_ms(2), flush=True)
    C = np.linalg.qr(rng.standard_normal((n, n)))[0]
    T = ctx.eri_transform(C)
    print("mo ms", ctx.last_kernel_ms(4), "sum", float(T.sum()), flush=True)
else:
    cfgs = [("sym", {}), ("sym", {"FLUSH_CLEAN": "0"}), ("tma", {}), ("sym", {"TUNA_B200_PDL": "0"})]
    for tile in ():
        for st in (3, 4):
            for cps in (1, 2):
                cfgs.append(("tma", dict(TUNA_B200_JK_TILE_KB=str(tile), TUNA_B200_JK_STAGES=str(st), TUNA_B200_JK_CTAS_PER_SM=str(cps))))
    for name in ("n2_ccpvtz", "ne2_uhf_ccpvqz", "et100"):
        for k, extra in cfgs:
            env = dict(os.environ, TUNA_B200_STORED_KERNEL=k, **extra)
            print(extra, end=" ")
            r = subprocess.run(["timeout", "120", sys.executable, __file__, "child", name], env=env, capture_output=True, text=True)
            print(name, k, "|", " | ".join(r.stdout.strip().splitlines()[-1:]) if r.stdout.strip() else r.stderr[-400:], flush=True)
