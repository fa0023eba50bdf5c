"""Re-export of the synthetic-input generators (they live in the package so that bench.py's product arm does not have
to import anything from oracle/).  TEST INFRASTRUCTURE ONLY."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
import rtvc_b200  # noqa: E402,F401
from rtvc_b200.synth import *  # noqa: E402,F401,F403
from rtvc_b200.synth import (AUX_DIMS, COMPUTE_DIMS, FC_DIMS, FEAT_DIMS, HOP, PAD, RES_BLOCKS, RES_OUT_DIMS, RNN_DIMS,  # noqa: E402,F401
                             UPSAMPLE, make_state_dict, make_state_dict_gn, make_state_dict_rr, n_classes, prune_state_dict,
                             synthetic_mel)
