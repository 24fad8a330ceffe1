import os, subprocess, sys
cfgs = []
for gdiv in (4, 8, 16, 32):
    for spl in (128, 256, 512):
        cfgs.append(dict(TUNA_B200_G_DIV=str(gdiv), TUNA_B200_SMEM_PER_LANE=str(spl)))
for itb in (2048, 4096):
    cfgs.append(dict(TUNA_B200_IT_BUDGET=str(itb), TUNA_B200_S_BUDGET=str(itb)))
for own in ("1e5", "2e6"):
    cfgs.append(dict(TUNA_B200_OWN_LAUNCH_MIN=own))
for c in cfgs:
    env = dict(os.environ, **c)
    r = subprocess.run([sys.executable, "tools/gsweep.py", "child", "400"], env=env, capture_output=True, text=True)
    print(c, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-200:], flush=True)
