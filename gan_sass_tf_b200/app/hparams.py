"""Hyper-parameters and plugin registries.

Same names, defaults and registry semantics as the reference's
``app/hparams.py`` (constants :8-43, registries :59-119): module-level constants
in CAPS that every op reads at call time, and five dict registries filled by
``@register_*`` decorators and read by ``get_*()`` (unknown name -> ``KeyError``).
``HOP_SIZE`` is new: the reference inherits SciPy's default ``nperseg // 2``
(main.py:97); ``None`` keeps that, the BASELINE configs set ``FFT_SIZE // 4``.
"""

BATCH_SIZE = 8
MAX_N_SIGNAL = 3

# width of the packed spectrogram; no dataset re-install is needed when this
# changes any more - the STFT runs on the device per batch
FFT_SIZE = 256
HOP_SIZE = None          # None -> FFT_SIZE // 2 (SciPy's default noverlap)
SAMPLE_RATE = 16000

CHARSET_SIZE = 27

FLOATX = 'float32'
INTX = 'int32'

RELU_LEAKAGE = 0.3
EPS = 1e-7
DROPOUT_KEEP_PROB = 0.8
REG_SCALE = 1e-2
REG_TYPE = 'L2'

USE_ASR = False
SEPARATOR_TYPE = 'toy'
RECOGNIZER_TYPE = 'toy'
DISCRIMINATOR_TYPE = 'toy'
OPTIMIZER_TYPE = 'adam'
LR = 1e-5
LR_DECAY = None

DATASET_TYPE = 'toy'

CTC_DECODER_TYPE = 'greedy'

SUMMARY_DIR = './logs'
ASR_SUMMARY_DIR = './asr_logs'

CLS_REAL_SIGNAL = 0
CLS_REAL_NOISE = 1
CLS_FAKE_SIGNAL = 2

assert isinstance(DROPOUT_KEEP_PROB, float)
assert 0. < DROPOUT_KEEP_PROB <= 1.
assert isinstance(LR, float) and LR >= 0.


def hop_size():
    """Hop in samples: ``HOP_SIZE`` or SciPy's default ``FFT_SIZE // 2``."""
    return FFT_SIZE // 2 if HOP_SIZE is None else int(HOP_SIZE)


separator_registry = {}
recognizer_registry = {}
discriminator_registry = {}
ozer_registry = {}
dataset_registry = {}


def _register(registry):
    def register(name):
        def wrapper(obj):
            registry[name] = obj
            return obj
        return wrapper
    return register


register_separator = _register(separator_registry)
register_recognizer = _register(recognizer_registry)
register_discriminator = _register(discriminator_registry)
register_optimizer = _register(ozer_registry)
register_dataset = _register(dataset_registry)


def get_separator():
    return separator_registry[SEPARATOR_TYPE]


def get_recognizer():
    return recognizer_registry[RECOGNIZER_TYPE]


def get_discriminator():
    return discriminator_registry[DISCRIMINATOR_TYPE]


def get_optimizer():
    return ozer_registry[OPTIMIZER_TYPE]


def get_dataset():
    return dataset_registry[DATASET_TYPE]
