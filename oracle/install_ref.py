"""Copies the handful of reference files the CPU arm imports (the unmodified vocoder package + config) from
/root/reference into baseline/_ref/ (git-ignored, travels to the GPU box with the snapshot).

TEST / BENCH INFRASTRUCTURE ONLY.  Nothing is edited: oracle/ref_import.py stubs the four third-party imports the tree
needs (matplotlib, librosa, soundfile, the pybind11 module) at import time.  The reference has no setup.py / pyproject,
so `pip install --target baseline/_ref /root/reference` has nothing to build; a plain copy of the sources it would have
installed is the install (recorded in DESIGN.md section 6)."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
WANT = ["config", "vocoder"]


def install(src="/root/reference", dest=DEST, verbose=False):
    if not os.path.isfile(os.path.join(src, "vocoder", "models", "fatchord_version.py")):
        return os.path.isfile(os.path.join(dest, "vocoder", "models", "fatchord_version.py"))
    for top in WANT:
        for dirpath, dirnames, filenames in os.walk(os.path.join(src, top)):
            dirnames[:] = [d for d in dirnames if d not in ("__pycache__", "src", "build")]
            rel = os.path.relpath(dirpath, src)
            os.makedirs(os.path.join(dest, rel), exist_ok=True)
            for f in filenames:
                if f.endswith(".py") or f == "LICENSE.txt":
                    shutil.copy2(os.path.join(dirpath, f), os.path.join(dest, rel, f))
                    if verbose:
                        print("copied", os.path.join(rel, f))
    return True


if __name__ == "__main__":
    ok = install(verbose="-v" in sys.argv)
    print("baseline/_ref", "ready" if ok else "unavailable (no /root/reference here and no previous copy)")
