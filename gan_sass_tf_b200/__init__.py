"""gan_sass_tf_b200 - B200-native (sm_100a) spectral hot path of GAN_SASS_TF.

``gan_sass_tf_b200.app`` mirrors the reference's ``app`` package (hparams / ops /
modules / utils / datasets) for the STFT -> log-feature -> mask -> iSTFT path;
the arithmetic lives in ``csrc/`` behind the C ABI of ``include/gss_api.h``.
"""
__version__ = "0.1.0"
