"""Attribute-style config objects standing in for OmegaConf ``DictConfig`` (absent from the
image): attribute access, ``.get(key, default)``, nested dicts.  The reference's constructors
only use those two access patterns (SAST.py:34-44, sast_rnn.py:71-99), so a real ``DictConfig``
works just as well."""
from __future__ import annotations


class Config(dict):
    def __getattr__(self, key):
        try:
            val = self[key]
        except KeyError as exc:
            raise AttributeError(key) from exc
        if isinstance(val, dict) and not isinstance(val, Config):
            val = Config(val)
            self[key] = val
        return val

    def __setattr__(self, key, val):
        self[key] = val

    def get(self, key, default=None):
        return getattr(self, key) if key in self else default


def attention_config(partition_size, **overrides) -> Config:
    """``model.backbone.stage.attention`` of config/model/sast_yolox/default.yaml:21-37."""
    cfg = dict(use_torch_mha=False, partition_size=tuple(partition_size), dim_head=32, attention_bias=True,
               mlp_activation="gelu", mlp_gated=False, mlp_bias=True, mlp_ratio=4, drop_mlp=0, drop_path=0,
               ls_init_value=1e-5, enable_CB=False, AMP=2e-4, BOUNCE=1e-3)
    cfg.update(overrides)
    return Config(cfg)


def backbone_config(in_res_hw, embed_dim=64, partition_split_32=2, num_blocks=(1, 1, 1, 1), enable_masking=False,
                    **attention_overrides) -> Config:
    """``model.backbone`` as config/modifier.py:27-41 completes it: in_res_hw must already be a
    multiple of 32*partition_split_32 (Gen1 240x304 -> 256x320, 1 Mpx -> 384x640)."""
    mult = 32 * partition_split_32
    assert in_res_hw[0] % mult == 0 and in_res_hw[1] % mult == 0, f"in_res_hw {in_res_hw} must be a multiple of {mult}"
    part = (in_res_hw[0] // mult, in_res_hw[1] // mult)
    return Config(dict(
        name="SASTRNN", input_channels=20, enable_masking=enable_masking, partition_split_32=partition_split_32,
        embed_dim=embed_dim, dim_multiplier=[1, 2, 4, 8], num_blocks=list(num_blocks),
        T_max_chrono_init=[4, 8, 16, 32], stem=dict(patch_size=4), in_res_hw=list(in_res_hw),
        stage=dict(downsample=dict(type="patch", overlap=True, norm_affine=True),
                   attention=dict(attention_config(part, **attention_overrides)),
                   lstm=dict(dws_conv=False, dws_conv_only_hidden=True, dws_conv_kernel_size=3, drop_cell_update=0))))
