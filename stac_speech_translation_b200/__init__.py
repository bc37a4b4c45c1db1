"""stac_speech_translation_b200: B200-native (sm_100a) encoder-side inference path of STAC-ST.

Drop-in classes for the objects the reference's HyperPyYAML instantiates on its hot path
(``compute_features``, ``normalize``, ``CNN``, ``Transformer``, ``ctc_lin``, ``log_softmax``),
backed by hand-written CUDA kernels behind a C ABI (include/stac_b200.h).  No CPU fallback.
"""
from ._lib import StacB200Error, LIB_PATH  # noqa: F401
from .features import Fbank, InputNormalization  # noqa: F401
from .convolution import ConvolutionFrontEnd  # noqa: F401
from .transformer import TransformerMultiTask, EncoderWrapper  # noqa: F401
from .linear import Linear, LogSoftmax  # noqa: F401
from .augment import SpecAugment  # noqa: F401
from .losses import ctc_loss  # noqa: F401
from .pipeline import (  # noqa: F401
    HParams, MODEL_SIZES, build_modules, compute_forward, EncoderPipeline, GraphedPipeline, ctc_greedy_collapse,
)

__all__ = [
    "Fbank", "InputNormalization", "ConvolutionFrontEnd", "TransformerMultiTask", "EncoderWrapper",
    "Linear", "LogSoftmax", "HParams", "MODEL_SIZES", "build_modules", "compute_forward",
    "EncoderPipeline", "GraphedPipeline", "ctc_greedy_collapse", "StacB200Error", "SpecAugment", "ctc_loss",
]
