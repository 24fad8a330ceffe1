import os, subprocess, sys
cfgs = [dict(), dict(TUNA_B200_G_DIV="16"), dict(TUNA_B200_G_DIV="4"), dict(TUNA_B200_SMEM_PER_LANE="384"), dict(TUNA_B200_SMEM_PER_LANE="192"),
        dict(TUNA_B200_IT_BUDGET="3072", TUNA_B200_S_BUDGET="3072"), dict(TUNA_B200_IT_BUDGET="8192", TUNA_B200_S_BUDGET="8192"),
        dict(TUNA_B200_OWN_LAUNCH_MIN="5e6")]
for c in cfgs:
    env = dict(os.environ, **c)
    r = subprocess.run([sys.executable, "tools/gsweep.py", "child", "800"], env=env, capture_output=True, text=True)
    print(c, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-200:], flush=True)
