"""Drop-in for the reference's ``models/detection/recurrent_backbone/sast_rnn.py``."""
from sast_b200.backbone import (RNNDetector, RNNDetectorStage, SASTAttentionPairCl, PositionEmbeddingSine,  # noqa: F401
                                non_zero_ratio)
