"""Minimal stand-in for ``omegaconf`` (absent from this image) so that the
reference's modules import unmodified inside ``oracle/gen_golden.py``.
Test infrastructure only; never imported by the product package."""
from contextlib import contextmanager


class DictConfig(dict):
    """Attribute-style dict with ``.get``; nested dicts are wrapped on access."""

    def __getattr__(self, key):
        try:
            val = self[key]
        except KeyError as exc:
            raise AttributeError(key) from exc
        if isinstance(val, dict) and not isinstance(val, DictConfig):
            val = DictConfig(val)
            self[key] = val
        return val

    def __setattr__(self, key, val):
        self[key] = val

    def get(self, key, default=None):
        return getattr(self, key) if key in self else default


ListConfig = list


class OmegaConf:
    @staticmethod
    def to_container(cfg, resolve=True, throw_on_missing=False):
        return dict(cfg)

    @staticmethod
    def create(obj):
        return DictConfig(obj)


@contextmanager
def open_dict(cfg):
    yield cfg
