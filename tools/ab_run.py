"""A/B: time kernel variants (real-time-voice-cloning_b200/_variants/*.so) against each other in one process tree."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cases = [(3410, 341, 1), (1705, 170, 2), (1140, 114, 3), (853, 85, 4)]
variants = sys.argv[1:]
for v in variants:
    name, _, envs = v.partition(":")
    env = dict(os.environ, WRNN_B200_LIB=os.path.join(root, "real-time-voice-cloning_b200", "_variants", name + ".so"))
    for kv in filter(None, envs.split(",")):
        k, _, val = kv.partition("=")
        env[k] = val
    for tg, ov, sets in cases:
        if name == "base" and sets > 2:
            continue
        e = dict(env, WRNN_TC_SETS=str(sets))
        print("%-28s" % v, end=" ", flush=True)
        subprocess.run([sys.executable, os.path.join(root, "tools", "sets_sweep.py"), str(tg), str(ov)], env=e)
