"""sast_b200 -- the SAST scene-adaptive sparse-attention block as hand-written CUDA for
NVIDIA B200 (sm_100a), behind the reference's own Python module API.

    from sast_b200 import SAST_block, MS_WSA, RNNDetector, build_recurrent_backbone, YoloXDetector

Importing the package does not load the shared library; the first op call does, and raises if
``sast_b200/libsast_b200.so`` has not been built (``make -C sast_b200/csrc``)."""
from .sast import (SAST_block, MS_WSA, PositiveLinear, LayerScale, MLP, GLU, LazyCount, window_partition,
                   window_reverse, grid_partition, grid_reverse, get_score_index_2d21d,
                   get_score_index_with_padding)
from .backbone import (RNNDetector, RNNDetectorStage, SASTAttentionPairCl, PositionEmbeddingSine,
                       ConvDownsampling_Cf2Cl, DWSConvLSTM2d, non_zero_ratio, build_recurrent_backbone)
from .ops import PackedEvents, pack_events
from .config import Config, attention_config, backbone_config
from .yolox import YoloXDetector, YOLOPAFPN, YOLOXHead, postprocess, detector_config

__version__ = "0.1.0"
