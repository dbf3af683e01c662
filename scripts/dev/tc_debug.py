import os, sys, subprocess
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for dbg in (0, 1, 2, 4, 5, 7):
    env = dict(os.environ, SS_TC_DEBUG=str(dbg))
    out = subprocess.run([sys.executable, os.path.join(root, "scripts/profile_target.py")], capture_output=True, text=True, timeout=200, env=env)
    l = [x for x in out.stdout.splitlines() if x.startswith("mpc")]
    print("debug", dbg, l[0][l[0].index("[("):][:130] if l else out.stderr[-300:])
