"""Dense ERI fill: shell-quartet engine fill mode vs the per-AO-quartet kernel (TUNA_B200_FILL_ENGINE=0): kernel times and max |diff| (development aid)."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tuna_b200  # noqa: E402
from util import load_golden, context_for  # noqa: E402

for name in sys.argv[1:] or ["n2_ccpvtz", "ne2_uhf_ccpvqz", "et100"]:
    g = load_golden(name)
    ctx = context_for(g)
    out = {"name": name, "ncart": int(ctx.ncart) if hasattr(ctx, "ncart") else None}
    tensors = {}
    for mode in ("1", "0"):
        os.environ["TUNA_B200_FILL_ENGINE"] = mode
        ts = []
        for _ in range(5):
            ctx.eri_fill_cart(); ts.append(ctx.last_kernel_ms(0))
        tensors[mode] = ctx.eri_download(0)
        out["engine_ms" if mode == "1" else "per_quartet_ms"] = [round(t, 4) for t in ts]
    out["max_abs_diff"] = float(np.abs(tensors["1"] - tensors["0"]).max())
    out["max_abs"] = float(np.abs(tensors["0"]).max())
    print(json.dumps(out), flush=True)
