"""CUDA-graph execution of the backbone: replay == eager, stateless and with carried LSTM state."""
import pytest
import torch

import sast_b200
from oracle.golden_common import event_histogram
from sast_b200.config import backbone_config
from sast_b200.runner import GraphedBackbone

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _net():
    torch.manual_seed(0)
    net = sast_b200.build_recurrent_backbone(backbone_config((128, 192), embed_dim=32)).to(DEV).eval()
    with torch.no_grad():      # make the attention / MLP branch visible (LayerScale init is 1e-5)
        for n, p in net.named_parameters():
            if n.endswith("gamma"):
                p.fill_(0.5)
    return net


def test_graph_replay_matches_eager_stateless():
    net = _net()
    xs = [event_histogram(2, 20, 128, 192, d, seed=i).to(DEV) for i, d in enumerate((0.05, 0.005, 0.3))]
    run = GraphedBackbone(net, xs[0], recurrent=False)
    for x in xs:
        with torch.no_grad():
            f_ref, s_ref, p_ref = net(x, None)
        f, s, raw = run(x)
        for st in (1, 2, 3, 4):
            assert torch.equal(f[st], f_ref[st]), st
        assert run.counts(raw) == [int(p) for p in p_ref]


def test_graph_replay_matches_eager_streaming():
    """21-step recurrent sequence with the state carried on the device, then a reset (RNNStates semantics,
    modules/utils/detection.py:76-130)."""
    net = _net()
    xs = [event_histogram(2, 20, 128, 192, 0.02 + 0.01 * (i % 3), seed=10 + i).to(DEV) for i in range(21)]
    run = GraphedBackbone(net, xs[0], recurrent=True)
    for rep in range(2):
        run.reset_states()
        states = None
        for i, x in enumerate(xs):
            with torch.no_grad():
                f_ref, states, p_ref = net(x, states)
            f, s, raw = run(x)
            assert torch.equal(f[4], f_ref[4]) and torch.equal(f[1], f_ref[1]), (rep, i)
            assert run.counts(raw) == [int(p) for p in p_ref]
