"""seq2seq-vc hot path on B200: hand-written sm_100a kernels behind the reference's model surface.

Public surface (mirrors seq2seq_vc.models / seq2seq_vc.losses / bin.preprocess names):
    VTN, TransformerTTS, AASVC, Seq2SeqLoss, GuidedMultiHeadAttentionLoss, L1Loss, ForwardSumLoss, DurationPredictorLoss,
    viterbi_decode, logmelfilterbank, LengthRegulator
The native library (libs2svc_b200.so) is loaded lazily on first use; there is no CPU fallback.
"""
from ._lib import S2SError  # noqa: F401
from .vtn_engine import VTNEngine, default_hparams  # noqa: F401
from .aasvc_engine import AASVCEngine  # noqa: F401
from .api import VTN, TransformerTTS, Seq2SeqLoss, GuidedMultiHeadAttentionLoss, VTNTrainStep, viterbi_decode, logmelfilterbank  # noqa: F401
from .api import AASVC, AASVCTrainStep, FastSpeechVC, NARVCTrainStep, L1Loss, ForwardSumLoss, DurationPredictorLoss  # noqa: F401
from .api import DistributedDataParallel, LengthRegulator  # noqa: F401
from .api import Conv2dSubsampling2, Conv2dSubsampling6, Conv2dSubsampling8, FeatureStatistics, Spectrogram2Waveform, griffin_lim, logmel2linear  # noqa: F401

AR_VC_MODELS = [VTN]
NAR_VC_MODELS = [FastSpeechVC, AASVC]
AR_TTS_MODELS = [TransformerTTS]
