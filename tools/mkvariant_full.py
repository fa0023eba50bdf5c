"""Build a FULL library variant (every source recompiled with extra flags, e.g. constants of engine_internal.h):
python tools/mkvariant_full.py NAME -DFLAG=V ... -> real-time-voice-cloning_b200/_variants/NAME.so"""
import os, subprocess, sys, tempfile
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = os.path.join(root, "real-time-voice-cloning_b200")
name, flags = sys.argv[1], sys.argv[2:]
os.makedirs(os.path.join(pkg, "_variants"), exist_ok=True)
tmp = tempfile.mkdtemp()
base = ["/usr/local/cuda/bin/nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]
objs, procs = [], []
for f in sorted(os.listdir(os.path.join(pkg, "csrc"))):
    if f.endswith(".cu"):
        o = os.path.join(tmp, f[:-3] + ".o")
        objs.append(o)
        procs.append(subprocess.Popen(base + flags + ["-c", os.path.join(pkg, "csrc", f), "-o", o], stderr=subprocess.DEVNULL))
for p in procs:
    assert p.wait() == 0
out = os.path.join(pkg, "_variants", name + ".so")
subprocess.run(["/usr/local/cuda/bin/nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs, check=True)
print(out)
