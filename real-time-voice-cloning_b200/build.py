"""Builds csrc/*.cu into libwavernn_b200.so (in-tree, sm_100a only) with nvcc.

nvcc cross-compiles without a GPU.  The library links the static CUDA runtime, so it has no
dependency on torch or on a libcudart path; it travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libwavernn_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--ptxas-options=-v"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


STAMP = os.path.join(HERE, "build_stamp.json")


def source_digest():
    """sha256 over every source the library is built from (names + contents) and the compiler flags: the library is rebuilt when
    -- and only when -- this changes (mtimes do not survive a snapshot copy; a stale .so next to newer sources must not pass)."""
    import hashlib
    h = hashlib.sha256(" ".join(FLAGS).encode())
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h", ".inc")))
    deps.append(os.path.join(HERE, "..", "include", "wavernn_b200.h"))
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    try:
        import json
        return json.load(open(STAMP)).get("sources_sha256") != source_digest()
    except Exception:
        return True


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    log = []
    for src in sources():
        obj = src[:-3] + ".o"
        cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(r.stderr)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed for %s" % src)
        objs.append(obj)
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    with open(os.path.join(HERE, "build_ptxas.log"), "w") as f:
        f.write("\n".join(log))
    import json
    import time
    ver = subprocess.run([NVCC, "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-1:]
    with open(STAMP, "w") as f:       # travels with the .so: which sources / flags / compiler it was built from, and where
        json.dump({"sources_sha256": source_digest(), "flags": FLAGS, "nvcc": ver, "built_at": time.strftime("%Y-%m-%dT%H:%M:%S"),
                   "host": os.uname().nodename, "objects": [os.path.basename(o) for o in objs]}, f, indent=1)
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
