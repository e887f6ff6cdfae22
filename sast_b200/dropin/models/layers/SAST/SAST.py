"""Drop-in for the reference's ``models/layers/SAST/SAST.py`` (same public names)."""
from sast_b200.sast import (SAST_block, MS_WSA, PositiveLinear, get_score_index_2d21d,  # noqa: F401
                            get_score_index_with_padding)
from sast_b200.backbone import non_zero_ratio as _nzr


def get_non_zero_ratio(x):
    """ref: SAST.py:284-302 (unused duplicate of sast_rnn.non_zero_ratio): list of four [B,C] ratios."""
    r = _nzr(x)
    return [r[:, i] for i in range(4)]
