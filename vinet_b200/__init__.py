"""vinet_b200 — B200-native (sm_100a) implementation of the ViNet / AViNet video-saliency hot path.

Public surface mirrors the reference (samyak0210/ViNet): ``VideoSaliencyModel`` (model.py:72),
``VideoAudioSaliencyModel`` (model.py:191, incl. ``use_transformer=True``), ``VideoAudioSaliencyFusionModel`` (model.py:116),
``kldiv/cc/similarity/nss`` (loss.py) and ``loss_func/get_loss`` (utils.py); plus the drivers around the path:
``GraphedTrainStep`` / ``GraphedForward``, ``SlidingWindowSaliency`` (+ ``AudioTrack``), ``FramePreprocessor`` / ``audio_window``.  Everything computes through ``libvinet_b200.so`` (include/vinet_b200.h).
"""
from .loss import cc, get_loss, kldiv, loss_func, nss, similarity  # noqa: F401
from .graph import GraphedForward, GraphedTrainStep  # noqa: F401
from .inference import AudioTrack, SlidingWindowSaliency  # noqa: F401
from .model import VideoSaliencyModel  # noqa: F401
from .preprocess import FramePreprocessor, audio_window  # noqa: F401

try:  # AViNet lives in its own module so a ViNet-only user never touches the audio kernels
    from .avmodel import VideoAudioSaliencyFusionModel, VideoAudioSaliencyModel  # noqa: F401
except ImportError:  # pragma: no cover
    pass

__all__ = ["VideoSaliencyModel", "VideoAudioSaliencyModel", "VideoAudioSaliencyFusionModel", "kldiv", "cc", "similarity", "nss", "loss_func",
           "get_loss", "GraphedTrainStep", "GraphedForward", "SlidingWindowSaliency", "AudioTrack", "FramePreprocessor", "audio_window"]
