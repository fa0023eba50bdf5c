"""B200-native WaveRNN vocoder inference engine (drop-in for the vocoder inference path of
RuntimeRacer/Real-Time-Voice-Cloning: vocoder.inference.infer_waveform -> WaveRNN.generate).

Import as `rtvc_b200` (the directory name of this package contains hyphens; rtvc_b200.py at the repo
root aliases it).  Layout mirrors the reference for the path only:
    rtvc_b200.vocoder.inference            load_model / is_loaded / infer_waveform / set_seed
    rtvc_b200.vocoder.models.base          type constants + init_voc_model
    rtvc_b200.vocoder.models.fatchord_version.WaveRNN   generate / fold_with_overlap / xfade_and_unfold / ...
    rtvc_b200.config.hparams               sp, wavernn_fatchord
    rtvc_b200._native                      ctypes binding of include/wavernn_b200.h
All numerics run in hand-written sm_100a CUDA behind the C ABI; there is no CPU fallback.
"""
__version__ = "0.1.0"
