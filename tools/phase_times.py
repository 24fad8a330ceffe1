"""Timing with individual phases of the shell engine skipped (development aid; results are wrong by construction)."""
import os, subprocess, sys
names = {0: "full", 1: "no boys(0)", 2: "no R/XY(1)", 4: "no U(2)", 8: "no S(3)", 16: "no assembly(4)", 32: "no digestion(5)", 64: "no staging", 128: "no flush",
         255: "nothing (decode+barriers only)", 48: "no 4+5", 14: "no 1+2+3"}
for mask, name in names.items():
    env = dict(os.environ, TUNA_B200_DBG_SKIP=str(mask))
    r = subprocess.run([sys.executable, "tools/gsweep.py", "child", sys.argv[1] if len(sys.argv) > 1 else "400"], env=env, capture_output=True, text=True)
    print(f"{mask:4d} {name:32s}", r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-200:], flush=True)
