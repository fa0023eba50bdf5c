"""Import the UNMODIFIED reference (RuntimeRacer/Real-Time-Voice-Cloning) vocoder path on CPU.

TEST INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py (in the build container, where
/root/reference exists) to mint golden vectors, and by bench.py --impl reference when a copy of
the needed reference files travels under baseline/_ref/.  Never imported by the product package.

The reference cannot be imported bare: vocoder/display.py:1 needs matplotlib,
vocoder/audio.py:3,6 need librosa + soundfile, vocoder/inference.py:3 pulls
vocoder/libwavernn/inference.py:11 which needs the pybind11 module WaveRNNVocoder, and
fatchord_version.py:64 uses np.cumproduct (removed in numpy 2).  We stub exactly those.
"""
import os
import sys
import types


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules.setdefault(name, m)
    return sys.modules[name]


def find_reference_root():
    here = os.path.dirname(os.path.abspath(__file__))
    cands = [os.environ.get("RTVC_REFERENCE_ROOT", ""), "/root/reference",
             os.path.join(here, "..", "baseline", "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "vocoder", "models", "fatchord_version.py")):
            return os.path.abspath(c)
    return None


def import_reference(root=None):
    """Returns (base_module, fatchord_module, hparams_module, inference_module)."""
    import numpy as np
    root = root or find_reference_root()
    if root is None:
        raise ImportError("reference tree not found (set RTVC_REFERENCE_ROOT)")
    if not hasattr(np, "cumproduct"):
        np.cumproduct = np.cumprod
    plt = _stub("matplotlib.pyplot")
    mpl = _stub("matplotlib", pyplot=plt)
    mpl.use = lambda *a, **k: None
    filt = _stub("librosa.filters")
    _stub("librosa", filters=filt)
    _stub("soundfile")
    _stub("WaveRNNVocoder")
    try:
        import psutil  # noqa: F401
    except Exception:
        _stub("psutil", cpu_count=lambda logical=True: os.cpu_count())
    if root not in sys.path:
        sys.path.insert(0, root)
    from vocoder.models import base, fatchord_version
    from config import hparams
    import vocoder.inference as inference
    return base, fatchord_version, hparams, inference
