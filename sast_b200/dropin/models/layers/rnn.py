"""Drop-in for the reference's ``models/layers/rnn.py``."""
from sast_b200.backbone import DWSConvLSTM2d  # noqa: F401
