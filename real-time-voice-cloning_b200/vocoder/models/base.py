"""Model factory and type constants (reference: vocoder/models/base.py:9-15 constants, :18-109
init_voc_model, :112-120 get_model_type)."""
from ...config.hparams import sp, wavernn_fatchord, wavernn_geneing, wavernn_runtimeracer
from .fatchord_version import WaveRNN as WaveRNNFatchord
from .geneing_version import WaveRNN as WaveRNNGeneing
from .runtimeracer_version import WaveRNN as WaveRNNRuntimeRacer

# Vocoder types (base.py:9-10) plus the backend this package adds next to them
VOC_TYPE_CPP = 'libwavernn'
VOC_TYPE_PYTORCH = 'pytorch'
VOC_TYPE_B200 = 'b200'

# Vocoder models (base.py:13-15)
MODEL_TYPE_FATCHORD = 'fatchord-wavernn'
MODEL_TYPE_GENEING = 'geneing-wavernn'
MODEL_TYPE_RUNTIMERACER = 'runtimeracer-wavernn'


def init_voc_model(model_type, device, override_hp_fatchord=None, override_hp_geneing=None,
                   override_hp_runtimeracer=None):
    """Same call shape and return value (model, pruner) as base.py:18.  `device` is a CUDA device index,
    a torch.device, or a string like "cuda:1".  Pruning is a training-time concern: pruner is None."""
    if model_type == MODEL_TYPE_RUNTIMERACER:                     # base.py:82-104
        hparams = override_hp_runtimeracer if override_hp_runtimeracer is not None else wavernn_runtimeracer
        cls = WaveRNNRuntimeRacer
    elif model_type == MODEL_TYPE_FATCHORD:
        hparams = override_hp_fatchord if override_hp_fatchord is not None else wavernn_fatchord
        cls = WaveRNNFatchord
    elif model_type == MODEL_TYPE_GENEING:                        # base.py:57-80
        hparams = override_hp_geneing if override_hp_geneing is not None else wavernn_geneing
        cls = WaveRNNGeneing
    else:
        raise NotImplementedError("Invalid model of type '%s' provided. Aborting..." % model_type)
    prod = 1
    for f in hparams.upsample_factors:
        prod *= f
    assert prod == sp.hop_size                                   # base.py:27
    model = cls(
        rnn_dims=hparams.rnn_dims, fc_dims=hparams.fc_dims, bits=hparams.bits, pad=hparams.pad,
        upsample_factors=hparams.upsample_factors, feat_dims=sp.num_mels, compute_dims=hparams.compute_dims,
        res_out_dims=hparams.res_out_dims, res_blocks=hparams.res_blocks, hop_length=sp.hop_size,
        sample_rate=sp.sample_rate, mode=hparams.mode, pruning=True, device=device)
    return model, None


def get_model_type(model):
    if isinstance(model, WaveRNNRuntimeRacer):
        return MODEL_TYPE_RUNTIMERACER
    if isinstance(model, WaveRNNGeneing):
        return MODEL_TYPE_GENEING
    if isinstance(model, WaveRNNFatchord):
        return MODEL_TYPE_FATCHORD
    raise NotImplementedError("Provided object is not a valid vocoder model.")
