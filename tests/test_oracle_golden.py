"""The oracle (oracle/sast_oracle.py) against outputs of the unmodified reference
(tests/golden/, produced by oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import sast_oracle as O
from oracle.golden_common import event_histogram, make_params

FTOL = 5e-6   # fp32 GEMM blocking differs with the host thread count; index sets stay exact


@pytest.fixture(autouse=True)
def _four_threads():
    n = torch.get_num_threads()
    torch.set_num_threads(4)   # what oracle/gen_golden.py used
    yield
    torch.set_num_threads(n)


BLOCKS = ["block_c64_w6x10", "block_c128_w8x10_b1", "block_c64_cb", "block_c64_dense", "block_c256_w3x5"]


def test_select_kat(golden):
    g = golden("select_kat")
    for ci, case in enumerate(g.meta):
        pw, pt = g.t(f"c{ci}_pw"), g.t(f"c{ci}_pt")
        iw = O.select_windows_from_probs(pw, 1 / case["N"], case["bounce"])
        assert torch.equal(iw, g.t(f"c{ci}_iw"))
        it, asy, K = O.select_tokens_from_probs(pt, 1 / case["T"], case["bounce"])
        assert torch.equal(asy, g.t(f"c{ci}_asy"))
        assert torch.equal(K, g.t(f"c{ci}_K"))
        assert torch.equal(it, g.t(f"c{ci}_it"))  # same torch.topk, same order


def test_nonzero_ratio(golden):
    g = golden("small_fns")
    for nm in ("u8", "i32", "f32"):
        assert torch.equal(O.non_zero_ratio(g.t(f"nzr_{nm}_x")), g.t(f"nzr_{nm}_r"))


def test_position_table(golden):
    g = golden("small_fns")
    assert torch.equal(O.position_table(12, 20, 64), g.t("pos_12_20_64"))
    assert torch.equal(O.position_table(12, 20, 64)[:6, :10], g.t("pos_12_20_64_slice"))
    assert torch.equal(O.position_table(8, 10, 128), g.t("pos_8_10_128"))


def test_partition_maps(golden):
    g = golden("small_fns")
    ids = torch.arange(2 * 12 * 20, dtype=torch.float32).view(2, 12, 20, 1)
    assert torch.equal(O.window_partition(ids, (6, 10)).reshape(-1).int(), g.t("win_ids_6x10"))
    assert torch.equal(O.grid_partition(ids, (6, 10)).reshape(-1).int(), g.t("grid_ids_6x10"))
    w = O.window_partition(ids, (6, 10))
    assert torch.equal(O.window_reverse(w, (6, 10), (12, 20)), ids)
    gr = O.grid_partition(ids, (6, 10))
    assert torch.equal(O.grid_reverse(gr, (6, 10), (12, 20)), ids)
    ids = torch.arange(16 * 30, dtype=torch.float32).view(1, 16, 30, 1)
    assert torch.equal(O.window_partition(ids, (8, 10)).reshape(-1).int(), g.t("win_ids_8x10"))
    assert torch.equal(O.grid_partition(ids, (8, 10)).reshape(-1).int(), g.t("grid_ids_8x10"))


@pytest.mark.parametrize("name", BLOCKS)
def test_block(golden, name):
    g = golden(name)
    m = g.meta
    params = make_params(m["shapes"], seed=m["seed"])
    pos = O.position_table(m["H"], m["W"], m["C"])
    y, cnt, lists = O.sast_block(g.t("x"), pos, g.t("r"), params, tuple(m["part"]), amp=m["AMP"],
                                 bounce=m["BOUNCE"], enable_CB=m["enable_CB"])
    assert cnt == int(g.arrays["count"])
    for li in range(2):
        for j, nm in enumerate(("iw", "it", "pad", "asy", "K")):
            assert torch.equal(lists[li][j], g.t(f"l{li}_{nm}")), (li, nm)
    assert (y - g.t("y")).abs().max().item() <= FTOL
    if "y2" in g:
        params2 = make_params(m["shapes2"], seed=m["seed2"])
        y2, cnt2, _ = O.sast_block(y, pos, g.t("r"), params2, tuple(m["part"]), first_block=False,
                                   index_list=lists, amp=m["AMP"], enable_CB=m["enable_CB"])
        assert cnt2 == int(g.arrays["count2"])
        assert (y2 - g.t("y2")).abs().max().item() <= FTOL


@pytest.mark.parametrize("name", BLOCKS)
def test_dense_equivalent(golden, name):
    """The dense-equivalent statement of a layer (no index lists) agrees with the
    sparse form to fp32 round-off: selection sets are all the layer needs."""
    g = golden(name)
    m = g.meta
    part = tuple(m["part"])
    T = part[0] * part[1]
    params = make_params(m["shapes"], seed=m["seed"])
    pos = O.position_table(m["H"], m["W"], m["C"])
    B, H, W = m["B"], m["H"], m["W"]
    N = H * W // T
    xw, _ = O.scoring(g.t("x"), pos, g.t("r"), params, part, m["AMP"])

    def mask(li):
        sel = torch.zeros(B * N, T, dtype=torch.bool)
        iw, asy = g.t(f"l{li}_iw"), g.t(f"l{li}_asy")
        sel.view(-1)[iw[asy // T] * T + asy % T] = True
        return sel

    x1 = O.ms_wsa_dense(xw, mask(0), O.sub(params, "win_attn"), B, m["enable_CB"])
    x1 = O.window_reverse(x1, part, (H, W))
    x2 = O.ms_wsa_dense(O.grid_partition(x1, part).reshape(B * N, T, -1), mask(1),
                        O.sub(params, "grid_attn"), B, m["enable_CB"])
    y = O.grid_reverse(x2, part, (H, W))
    assert (y - g.t("y")).abs().max().item() < 2e-5


@pytest.mark.parametrize("name", ["backbone_e32", "backbone_e32_nb2_mask_cb"])
def test_backbone(golden, name):
    g = golden(name)
    m = g.meta
    params = make_params(m["shapes"], seed=m["seed"])
    cfg = dict(embed_dim=m["embed_dim"], dim_multiplier=[1, 2, 4, 8], num_blocks=m["num_blocks"],
               patch_size=4, in_res_hw=m["in_res_hw"], partition_size=m["partition_size"],
               AMP=m["AMP"], BOUNCE=m["BOUNCE"], enable_CB=m["enable_CB"])
    H, W = m["in_res_hw"]
    x0 = event_histogram(m["B"], 20, H, W, m["x_density"][0], seed=m["x_seeds"][0])
    x1 = event_histogram(m["B"], 20, H, W, m["x_density"][1], seed=m["x_seeds"][1])
    tm = g.t("token_mask") if "token_mask" in g else None
    f0, s0, p0 = O.backbone_forward(x0, None, params, cfg, tm)
    f1, s1, p1 = O.backbone_forward(x1, s0, params, cfg, tm)
    assert list(p0) == g.arrays["P0"].tolist() and list(p1) == g.arrays["P1"].tolist()
    for st in (1, 2, 3, 4):
        h = f1[st][:, :, ::2, ::2] if st == 1 else f1[st]
        assert (h - g.t(f"h1_s{st}")).abs().max().item() <= FTOL, st
        np.testing.assert_allclose([f0[st].double().sum().item(), f0[st].double().abs().sum().item()],
                                   g.arrays[f"sum0_s{st}"], rtol=1e-6)
        np.testing.assert_allclose([s1[st - 1][1].double().sum().item(), s1[st - 1][1].double().abs().sum().item()],
                                   g.arrays[f"csum1_s{st}"], rtol=1e-6)
