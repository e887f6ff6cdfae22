"""Drop-in for the reference's ``models/detection/recurrent_backbone/base.py``."""
from sast_b200.backbone import BaseDetector  # noqa: F401
