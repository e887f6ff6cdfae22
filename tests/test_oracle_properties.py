"""Size-independent properties of the oracle (CPU, seeded): the invariants the CUDA path is built on.
The golden fixtures pin values; these pin structure -- index conventions of SURVEY.md appendix A,
the dense-equivalent statement of a layer (section 8a), partition round trips."""
import pytest
import torch

from oracle import sast_oracle as O
from oracle.golden_common import make_params


@pytest.mark.parametrize("seed,B,N,T,bounce", [(0, 2, 16, 60, 1e-3), (1, 1, 4, 80, 1e-3), (2, 3, 64, 15, 0.5), (3, 1, 1, 80, 1e-3)])
def test_selection_index_conventions(seed, B, N, T, bounce):
    g = torch.Generator().manual_seed(seed)
    C = 8
    scores = torch.rand(B, N, T, C, generator=g) * torch.rand(B, N, 1, 1, generator=g) * 3
    iw, it, pad, asy, K = O.select_layer(scores, T, bounce)
    M = iw.numel()
    # index_window: ascending flat ids b*N+n; every frame keeps at least its arg-max window
    assert torch.equal(iw, iw.sort()[0]) and iw.unique().numel() == M and M >= B
    assert 0 <= iw.min() and iw.max() < B * N
    assert set((iw // N).tolist()) == set(range(B))
    # K: per selected window, >= 1 (softmax max >= 1/T > threshold), asy_index ascending in the compacted [M*T] space
    assert K.shape == (M,) and K.min() >= 1 and K.max() <= T and int(K.sum()) == asy.numel()
    assert torch.equal(asy, asy.sort()[0]) and asy.unique().numel() == asy.numel()
    assert torch.equal(torch.bincount(asy // T, minlength=M), K)
    # index_token: Kmax entries per window, inside that window, superset of asy_index; padding = the rest
    Kmax = int(K.max())
    assert it.numel() == M * Kmax
    assert torch.equal(it.view(M, Kmax) // T, torch.arange(M).view(-1, 1).expand(M, Kmax))
    assert torch.isin(asy, it).all()
    assert pad.numel() == it.numel() - asy.numel() and not torch.isin(pad, asy).any()
    if N == 1:
        assert M == B                     # a single window has probability 1.0: always selected


def test_selection_is_monotone_in_bounce():
    """A larger BOUNCE lowers the threshold d/(1+b): it never drops a window or a token."""
    g = torch.Generator().manual_seed(7)
    pw = torch.rand(2, 32, generator=g).softmax(-1)
    pt = torch.rand(9, 60, generator=g).softmax(-1)
    prev_w, prev_t = None, None
    for b in (0.0, 1e-3, 0.1, 1.0, 10.0):
        w = set(O.select_windows_from_probs(pw, 1 / 32, b).tolist())
        t = set(O.select_tokens_from_probs(pt, 1 / 60, b)[1].tolist())
        if prev_w is not None:
            assert prev_w <= w and prev_t <= t
        prev_w, prev_t = w, t


def test_threshold_compare_is_fp32():
    """`x >= d/(1+b)` casts the Python double to fp32 first: for T = 80 the fp32 value of the threshold is below
    the double, and an element equal to it in fp32 is still selected."""
    thr = (1 / 80) / (1 + 1e-3)
    t32 = torch.tensor(thr, dtype=torch.float32)
    assert float(t32) < thr
    prob = torch.full((1, 80), 0.0)
    prob[0, 3] = t32
    prob[0, 5] = torch.nextafter(t32, torch.tensor(0.0))
    _, asy, K = O.select_tokens_from_probs(prob, 1 / 80, 1e-3)
    assert asy.tolist() == [3] and K.tolist() == [1]


@pytest.mark.parametrize("part,hw", [((6, 10), (24, 40)), ((8, 10), (16, 20)), ((3, 5), (12, 20))])
def test_partition_round_trips(part, hw):
    x = torch.arange(2 * hw[0] * hw[1] * 3, dtype=torch.float32).view(2, hw[0], hw[1], 3)
    assert torch.equal(O.window_reverse(O.window_partition(x, part), part, hw), x)
    assert torch.equal(O.grid_reverse(O.grid_partition(x, part), part, hw), x)
    w = O.window_partition(x, part)
    g = O.grid_partition(x, part)
    assert w.shape == g.shape == (2 * hw[0] * hw[1] // (part[0] * part[1]), part[0], part[1], 3)
    # a window is a contiguous patch, a grid window a stride-(H/g0, W/g1) lattice
    assert torch.equal(w[1], x[0, :part[0], part[1]:2 * part[1]])
    assert torch.equal(g[1], x[0, 0::hw[0] // part[0], 1::hw[1] // part[1]])


@pytest.mark.parametrize("seed,C,part,B,hw,cb", [(0, 64, (6, 10), 2, (12, 20), False), (1, 32, (3, 5), 1, (6, 10), True)])
def test_sparse_layer_equals_dense_statement(seed, C, part, B, hw, cb):
    """ms_wsa (index lists, top-k padding, -1e4 column mask, as the reference) == the dense-equivalent statement
    (per-token mask, keys masked out): padding and top-k are implementation artefacts."""
    torch.manual_seed(seed)
    T = part[0] * part[1]
    N = hw[0] * hw[1] // T
    I = (int(C * 4) * 2 // 3 // 32) * 32
    shapes = {"norm1.weight": (C,), "norm1.bias": (C,), "norm2.weight": (C,), "norm2.bias": (C,), "qkv.weight": (3 * C, C),
              "qkv.bias": (3 * C,), "proj.weight": (C, C), "proj.bias": (C,), "ls1.gamma": (C,), "ls2.gamma": (C,),
              "mlp.net.0.proj.weight": (2 * I, C), "mlp.net.0.proj.bias": (2 * I,), "mlp.net.2.weight": (C, I), "mlp.net.2.bias": (C,)}
    p = make_params(shapes, seed=seed + 10)
    x = torch.randn(B * N, T, C)
    scores = torch.rand(B, N, T, 4) * torch.rand(B, N, 1, 1) * 2
    iw, it, pad, asy, K = O.select_layer(scores, T, 1e-3)
    y = O.ms_wsa(x, iw, it, pad, asy, iw.numel(), B, cb, p)
    sel = torch.zeros(B * N, T, dtype=torch.bool)
    sel[iw[asy // T], asy % T] = True
    assert 0 < sel.sum() < sel.numel()
    yd = O.ms_wsa_dense(x, sel, p, B, enable_CB=cb)
    assert y.shape == yd.shape and (y - yd).abs().max() < 2e-5
    # unselected tokens keep norm1(x), not x
    n1 = torch.nn.functional.layer_norm(x, (C,), p["norm1.weight"], p["norm1.bias"], 1e-5)
    assert torch.allclose(y[~sel], n1[~sel], atol=1e-6)
