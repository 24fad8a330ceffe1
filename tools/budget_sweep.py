import os, subprocess, sys
for itb, sb, spl in ((6144, 6144, 256), (3072, 3072, 256), (2048, 2048, 256), (1024, 2048, 256), (1024, 1024, 256), (2048, 2048, 128), (1024, 1024, 128), (512, 1024, 128), (2048, 2048, 512)):
    env = dict(os.environ, TUNA_B200_IT_BUDGET=str(itb), TUNA_B200_S_BUDGET=str(sb), TUNA_B200_SMEM_PER_LANE=str(spl))
    r = subprocess.run([sys.executable, "tools/gsweep.py", "child", "400"], env=env, capture_output=True, text=True)
    print(itb, sb, spl, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-200:], flush=True)
