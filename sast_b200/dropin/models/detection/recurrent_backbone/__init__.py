"""Drop-in for the reference's ``models/detection/recurrent_backbone`` package."""
from sast_b200.backbone import build_recurrent_backbone  # noqa: F401
from sast_b200.backbone import RNNDetector as SASTRNNDetector  # noqa: F401
