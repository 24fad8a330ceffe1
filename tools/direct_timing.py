"""A/B timing of development builds of the shell engine (build/*.so, selected with TUNA_B200_LIB) and of its runtime knobs.
Development aid: each child prints ms per direct build and two weighted checksums of J and K for a quick parity cross-check."""
import json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import tuna_b200
    from tuna_b200 import workloads as w
    from tuna_b200.basis import flatten, from_arrays
    nbf = int(sys.argv[2])
    b = w.even_tempered_diatomic(nbf)
    bfs = from_arrays(b["origins"], b["lmn"], b["nprim"], b["exps"], b["raw_coefs"])
    ctx = tuna_b200.Context(0)
    ctx.set_basis(*flatten(bfs)); ctx.set_transform(np.eye(len(bfs)))
    P = w.fixed_density(len(bfs))
    best = 1e30
    for _ in range(4):          # with TUNA_B200_GRAPH=1: plain, captured, replayed, replayed
        J, K = ctx.jk_direct(P, 1e-16)
        best = min(best, ctx.last_kernel_ms(3))
    W = np.random.default_rng(5).standard_normal(J.shape[-2:])
    print(json.dumps({"ms": best, "cj": float((np.asarray(J).reshape(W.shape) * W).sum()), "ck": float((np.asarray(K).reshape(W.shape) * W).sum()),
                      "nj": float(np.abs(J).max()), "nk": float(np.abs(K).max())}))
else:
    sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [400, 800]
    # build/*.so come from tools/build_variants.sh (compile-time variants of the shell engine, off in the shipped library)
    libs = {"cur": "tuna_b200/libtuna_b200.so", "wide": "build/lib_wide.so", "asm": "build/lib_asm.so", "tiers": "build/lib_tiers.so", "v3": "build/lib_v3.so"}
    combos = [("cur", {}), ("cur", {"TUNA_B200_GRAPH": "1"}), ("wide", {}), ("asm", {}), ("tiers", {}), ("tiers", {"TUNA_B200_REG_TIER": "0"}), ("v3", {}), ("v3", {"TUNA_B200_REG_TIER": "0"})]
    ref = {}
    for nbf in sizes:
        for name, env in combos:
            lib = libs[name]
            if not os.path.exists(os.path.join(ROOT, lib)):
                continue
            if True:
                e = dict(os.environ, TUNA_B200_LIB=os.path.join(ROOT, lib), **env)
                r = subprocess.run([sys.executable, __file__, "child", str(nbf)], env=e, capture_output=True, text=True)
                out = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "ERR " + r.stderr[-300:]
                note = ""
                try:
                    d = json.loads(out)
                    if nbf not in ref:
                        ref[nbf] = d
                    note = " dJ=%.2e dK=%.2e" % (abs(d["cj"] - ref[nbf]["cj"]) / ref[nbf]["nj"], abs(d["ck"] - ref[nbf]["ck"]) / ref[nbf]["nk"])
                except Exception:
                    pass
                print(nbf, name, env, out, note, flush=True)
