"""WaveRNN (runtimeracer topology: four GRU-256 cells on a residual chain, five FC layers) backed by the B200 engine.

Mirrors the inference surface of the reference class (vocoder/models/runtimeracer_version.py:97-134 constructor and layers,
:199-314 generate, fold / crossfade helpers shared with the fatchord class) -- the model type `vocoder/inference.py:43`
hard-codes for the reference's C++ path.  Same native engine and C ABI as the fatchord class (include/wavernn_b200.h,
`wrnn_set_topology(WRNN_TOPO_RUNTIMERACER)`): the front end, the fold plan and the post chain are shared, the sample loop
is `wrnn_loop_rr_kernel` (csrc/loop_rr.cu, fp32).  state_dict names are the reference's: I, rnn1..rnn4, fc1..fc5, upsample.*.
"""
from ... import _native
from .fatchord_version import WaveRNN as _WaveRNNBase

_FIXED = dict(rnn_dims=256, fc_dims=256, pad=2, upsample_factors=(5, 5, 8), feat_dims=80, compute_dims=128,
              res_out_dims=128, res_blocks=10, hop_length=200)


class WaveRNN(_WaveRNNBase):
    _FIXED_DIMS = _FIXED
    _TOPOLOGY = _native.TOPO_RUNTIMERACER
