"""Timing of the dense ERI fill (both engines) and of the Cartesian->spherical rotation (development aid)."""
import json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import tuna_b200
    from util import load_golden, context_for
    g = load_golden(sys.argv[2])
    ctx = context_for(g); ctx.set_transform(g["U"])
    tf, tr = [], []
    for _ in range(4):
        ctx.eri_fill_cart(); tf.append(ctx.last_kernel_ms(0))
        ctx.eri_cart_to_sph(); tr.append(ctx.last_kernel_ms(1))
    print(json.dumps({"name": sys.argv[2], "engine": os.environ.get("TUNA_B200_FILL_ENGINE", "default"), "fill_ms": min(tf), "rot_ms": min(tr), "fill_all": tf, "rot_all": tr}))
else:
    for name in ("n2_ccpvtz", "et100", "ne2_uhf_ccpvqz"):
        for eng in ("generic", "shell"):
            r = subprocess.run(["timeout", "120", sys.executable, __file__, "child", name], env=dict(os.environ, TUNA_B200_FILL_ENGINE=eng), capture_output=True, text=True)
            print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "ERR " + r.stderr[-300:], flush=True)
