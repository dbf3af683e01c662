import os, sys, shutil, subprocess
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
libs = sorted(os.listdir(os.path.join(root, "scripts/dev/libs")))
for lib in libs:
    shutil.copy(os.path.join(root, "scripts/dev/libs", lib), os.path.join(root, "smartstartcontinuous_b200/libss_b200.so"))
    out = subprocess.run([sys.executable, os.path.join(root, "scripts/dev/tc_check.py")], capture_output=True, text=True, timeout=200)
    lines = [l for l in out.stdout.splitlines() if "bf16_tc" in l and ("MPC" in l) and "per_sample" in l] + [l for l in out.stdout.splitlines() if "score err" in l][2:3]
    print(lib); [print("   ", l[:150]) for l in lines]
    if out.returncode: print(out.stderr[-500:])
