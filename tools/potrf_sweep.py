import sys, os, subprocess
for nb in sys.argv[2:]:
    env = dict(os.environ, PB_POTRF_NB=nb)
    for n in sys.argv[1].split(","):
        out = subprocess.run([sys.executable, "tools/potrf_once.py", n, "2"], env=env, capture_output=True, text=True).stdout.strip().splitlines()
        print("NB", nb, out[-1])
