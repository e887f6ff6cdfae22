"""Launch-overhead-free execution of the recurrent backbone: the whole 4-stage forward
(conv stems, SAST blocks, LSTMs) captured once in a CUDA graph and replayed per step.

Possible because the SAST block keeps every data-dependent count on the device (no host
sync between kernels).  Covers both protocols of the reference:
  * ``benchmark.py``-style stateless forward (``previous_states=None`` every call), and
  * streaming inference with LSTM states carried on the device across steps
    (modules/utils/detection.py:76-130 ``RNNStates`` semantics: reset -> zero states)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

Tensor = torch.Tensor


class GraphedBackbone:
    def __init__(self, net, example_x: Tensor, recurrent: bool = False, warmup: int = 3):
        assert example_x.is_cuda, "GraphedBackbone needs CUDA tensors"     # a Tensor or a sast_b200.PackedEvents
        self.net = net
        self.recurrent = recurrent
        self.x = example_x.clone()
        self.graph = torch.cuda.CUDAGraph()
        self.states: Optional[List[Tuple[Tensor, Tensor]]] = None
        side = torch.cuda.Stream(device=example_x.device)
        side.wait_stream(torch.cuda.current_stream(example_x.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                feats, states, P = net(self.x, None)
            if recurrent:
                self.states = [(torch.zeros_like(h), torch.zeros_like(c)) for h, c in states]
                for _ in range(2):
                    feats, states, P = net(self.x, self.states)
        torch.cuda.current_stream(example_x.device).wait_stream(side)
        torch.cuda.synchronize(example_x.device)
        with torch.cuda.graph(self.graph), torch.no_grad():
            feats, states, P = net(self.x, self.states if recurrent else None)
            if recurrent:
                for (hs, cs), (h, c) in zip(self.states, states):
                    hs.copy_(h)
                    cs.copy_(c)
            # raw selected-token totals of every SAST layer (static buffers of the graph); see counts()
            self._terms = [p.terms for p in P]
            self.raw_counts = torch.stack([t.reshape(()) for terms in self._terms for t, _ in terms])
        self.feats = feats
        self.new_states = states

    def reset_states(self):
        if self.states is not None:
            for h, c in self.states:
                h.zero_()
                c.zero_()

    def __call__(self, x: Optional[Tensor] = None):
        """Replay on ``x`` (copied into the graph's static input).  Returns the static output
        tensors: {stage: h}, states, per-stage selected-token counts [4] (device int tensor)."""
        if x is not None and x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.feats, self.new_states, self.raw_counts

    def counts(self, raw: Optional[Tensor] = None) -> List[int]:
        """Per-stage selected-token counts (the reference's P list) from a raw_counts tensor (syncs)."""
        vals = (self.raw_counts if raw is None else raw).tolist()
        out, i = [], 0
        for terms in self._terms:
            out.append(sum(int(vals[i + j]) // d for j, (_, d) in enumerate(terms)))
            i += len(terms)
        return out
