"""A/B: time library variants (real-time-voice-cloning_b200/_variants/*.so) on given cases: python tools/ab_cases.py MODE:T:tg:ov[,..] variant[:ENV=V,..] ..."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cases = [c.split(":") for c in sys.argv[1].split(",")]
for v in sys.argv[2:]:
    name, _, envs = v.partition(":")
    env = dict(os.environ)
    if name != "tree":
        env["WRNN_B200_LIB"] = os.path.join(root, "real-time-voice-cloning_b200", "_variants", name + ".so")
    for kv in filter(None, envs.split(",")):
        k, _, val = kv.partition("=")
        env[k] = val
    for mode, T, tg, ov in cases:
        print("%-24s" % v, end=" ", flush=True)
        subprocess.run([sys.executable, os.path.join(root, "tools", "one_time.py"), mode, T, tg, ov], env=env)
