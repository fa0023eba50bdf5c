"""Build a library variant for A/B timing: python tools/mkvariant.py NAME [-DFLAG=V ...] -> real-time-voice-cloning_b200/_variants/NAME.so
(loop_rs.cu recompiled with the extra flags, linked with the tree's other objects; select it with WRNN_B200_LIB=<path>)."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = os.path.join(root, "real-time-voice-cloning_b200")
name, flags = sys.argv[1], sys.argv[2:]
os.makedirs(os.path.join(pkg, "_variants"), exist_ok=True)
obj = os.path.join(pkg, "_variants", name + "_loop_rs.o")
base = ["/usr/local/cuda/bin/nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]
subprocess.run(base + flags + ["-c", os.path.join(pkg, "csrc", "loop_rs.cu"), "-o", obj], check=True)
objs = [os.path.join(pkg, "csrc", f) for f in sorted(os.listdir(os.path.join(pkg, "csrc"))) if f.endswith(".o") and f != "loop_rs.o"]
out = os.path.join(pkg, "_variants", name + ".so")
subprocess.run(["/usr/local/cuda/bin/nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, obj] + objs, check=True)
print(out)
