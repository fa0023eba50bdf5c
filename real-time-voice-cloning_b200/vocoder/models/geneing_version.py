"""WaveRNN (geneing topology: one GRU-256, fc1 with ReLU, fc3; aux split in two) backed by the B200 engine.

Mirrors the inference surface of the reference class (vocoder/models/geneing_version.py:89-121 constructor and layers, :157-252
generate; fold / crossfade helpers shared with the fatchord class).  Same native engine and C ABI as the fatchord class
(include/wavernn_b200.h, `wrnn_set_topology(WRNN_TOPO_GENEING)`): the front end has this topology's sizes (64 channels, 3
residual blocks, upsampling 4 x 5 x 10 -- config/hparams.py:288-300), the sample loop is `wrnn_loop_gn_kernel`
(csrc/loop_gn.cu, fp32).  Modes: 'BITS' (softmax over 2**bits classes; sampled like the fatchord RAW mode) and 'MOL'; the
beta-distribution mode the reference calls 'RAW' here (geneing_version.py:96-97, :213-216) is not supported.
"""
from ... import _native
from .fatchord_version import WaveRNN as _WaveRNNBase

_FIXED = dict(rnn_dims=256, fc_dims=128, pad=2, upsample_factors=(4, 5, 10), feat_dims=80, compute_dims=64,
              res_out_dims=64, res_blocks=3, hop_length=200)


class WaveRNN(_WaveRNNBase):
    _FIXED_DIMS = _FIXED
    _TOPOLOGY = _native.TOPO_GENEING

    def __init__(self, rnn_dims, fc_dims, bits, pad, upsample_factors, feat_dims, compute_dims, res_out_dims,
                 res_blocks, hop_length, sample_rate, mode='BITS', pruning=False, device=0):
        if mode == 'RAW':
            raise NotImplementedError("geneing mode 'RAW' (beta distribution, geneing_version.py:213-216) is not built; use 'BITS' or 'MOL'")
        if mode not in ('BITS', 'MOL'):
            raise ValueError("input_type: %s not supported" % mode)          # geneing_version.py:104
        super().__init__(rnn_dims, fc_dims, bits, pad, upsample_factors, feat_dims, compute_dims, res_out_dims, res_blocks,
                         hop_length, sample_rate, mode='RAW' if mode == 'BITS' else mode, pruning=pruning, device=device)
        self.mode = mode
        self.aux_dims = res_out_dims // 2
