"""Drop-in for the reference's ``models/layers/SAST/ops.py`` (same public names)."""
from sast_b200.sast import (LayerScale, MLP, GLU, window_partition, window_reverse, grid_partition,  # noqa: F401
                            grid_reverse)
from sast_b200.backbone import ConvDownsampling_Cf2Cl, get_downsample_layer_Cf2Cl  # noqa: F401


def nChw_2_nhwC(x):
    assert x.ndim == 4
    return x.permute(0, 2, 3, 1).contiguous()


def nhwC_2_nChw(x):
    assert x.ndim == 4
    return x.permute(0, 3, 1, 2).contiguous()
