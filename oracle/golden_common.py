"""Shared by ``oracle/gen_golden.py`` (runs the real reference in the build
container) and by ``tests/`` (which replay the fixtures without the reference):
deterministic, platform-independent weights and inputs from numpy seeds.
Test infrastructure only."""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import numpy as np
import torch


def make_params(shapes: Dict[str, Sequence[int]], seed: int, gamma: float = 0.5) -> Dict[str, torch.Tensor]:
    """Draw one tensor per state-dict key from ``numpy.random.RandomState(seed)``.

    Keys are visited in sorted order so the draw does not depend on dict order.
    Aliased keys (the reference registers MS_WSA sub-modules twice, under their own
    name and under ``sub_layers.N``) must be resolved by the caller: pass only the
    canonical keys."""
    rng = np.random.RandomState(seed)
    out = {}
    for key in sorted(shapes):
        shape = tuple(shapes[key])
        z = rng.standard_normal(shape).astype(np.float32)
        leaf = key.split(".")[-1]
        parent = key.split(".")[-2] if "." in key else ""
        if leaf == "gamma":
            val = gamma + 0.1 * z
        elif leaf == "mask_token":
            val = 0.02 * z
        elif parent.startswith("norm") and leaf == "weight":
            val = 1.0 + 0.1 * z
        elif leaf == "bias":
            val = 0.1 * z
        elif parent == "to_controls":
            val = 0.5 + 0.3 * z
        else:
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            val = z / np.sqrt(fan_in)
        out[key] = torch.from_numpy(np.ascontiguousarray(val, dtype=np.float32))
    return out


def make_det_params(shapes: Dict[str, Sequence[int]], seed: int) -> Dict[str, torch.Tensor]:
    """Parameters and BatchNorm statistics of the YOLOX neck / head from ``numpy.random.RandomState(seed)``:
    conv weights with a gain that keeps activations O(1) through ~20 layers, BN scale around 1, positive
    running variances, predictor biases around 0 (so that post-processing sees a few hundred candidates)."""
    rng = np.random.RandomState(seed)
    out = {}
    for key in sorted(shapes):
        shape = tuple(shapes[key])
        leaf = key.split(".")[-1]
        parent = key.split(".")[-2] if "." in key else ""
        if leaf == "num_batches_tracked":
            out[key] = torch.zeros(shape, dtype=torch.int64)
            continue
        z = rng.standard_normal(shape).astype(np.float32)
        if parent == "bn" and leaf == "weight":
            val = 1.0 + 0.1 * z
        elif parent == "bn" and leaf == "bias":
            val = 0.1 * z
        elif leaf == "running_mean":
            val = 0.1 * z
        elif leaf == "running_var":
            val = 0.6 + 0.4 * np.abs(z)
        elif leaf == "bias":
            val = 0.5 * z
        else:
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            gain = 0.3 if "_preds." in key else 1.4          # predictors: box regressions of O(1) before the exp
            val = gain * z / np.sqrt(fan_in)
        out[key] = torch.from_numpy(np.ascontiguousarray(val, dtype=np.float32))
    return out


def canonical_keys(state_dict) -> Dict[str, Tuple[int, ...]]:
    """Shapes of the state dict without the ``sub_layers.*`` aliases."""
    return {k: tuple(v.shape) for k, v in state_dict.items() if ".sub_layers." not in k}


ALIASES = {"sub_layers.0": "ls1", "sub_layers.2": "norm2", "sub_layers.3": "mlp", "sub_layers.4": "ls2"}


def with_aliases(params: Dict[str, torch.Tensor], like_keys) -> Dict[str, torch.Tensor]:
    """Expand canonical params to every key in ``like_keys`` (adds sub_layers aliases)."""
    out = {}
    for k in like_keys:
        src = k
        for a, b in ALIASES.items():
            src = src.replace("." + a + ".", "." + b + ".")
            if src.startswith(a + "."):
                src = b + src[len(a):]
        out[k] = params[src]
    return out


def event_histogram(B: int, bins: int, H: int, W: int, density: float, seed: int,
                    dtype=np.uint8) -> torch.Tensor:
    """Synthetic stacked event histogram: Poisson counts clipped at 10 (the
    reference's ``count_cutoff``), non-zero on roughly ``density`` of the pixels,
    with a per-frame density gradient so frames differ in sparsity."""
    rng = np.random.RandomState(seed)
    x = np.zeros((B, bins, H, W), dtype=np.float32)
    for b in range(B):
        dens = density * (0.3 + 1.4 * (b + 1) / B)
        blob = rng.rand(bins, H, W) < dens
        # spatially non-uniform: events concentrated in one half-plane that moves with b
        ramp = np.linspace(0.0, 1.0, W, dtype=np.float32)[None, None, :]
        keep = rng.rand(bins, H, W) < (0.15 + 0.85 * np.roll(ramp, b * W // max(B, 1), axis=2))
        cnt = np.minimum(rng.poisson(2.0, size=(bins, H, W)) + 1, 10)
        x[b] = blob * keep * cnt
    return torch.from_numpy(x.astype(dtype))
