"""GPU parity of the input-space formulation of the first layer (include/gnnfd_b200.h section (5)): the same
reference stages (PyG GATConv.forward / backward at src/models/gat.py:39,80, src/train.py:142) evaluated without
projected features, through the drop-in module (gemm_algo=GEMM_INPUT) and the C ABI, against the fp64 CPU oracle.
Tolerances are BASELINE.json's (fp32 max-abs 1e-5) plus relative L2 <= 1e-5 against fp64."""
import numpy as np
import pytest
import torch

from gnn_fraud_detection_b200 import GATConv, _abi, build_csr, functional as Fn, synth
from oracle import pyg_gatconv as O
from test_gat_gpu import _assert_close, _fwd_bwd_case, _layer
from util import maxabs, relerr, seeded_params

pytestmark = pytest.mark.gpu
KEYS = ("out", "alpha", "dW", "datt_src", "datt_dst", "dbias")


@pytest.mark.parametrize("N,E,K", [(1, 0, 5), (2, 1, 3), (64, 0, 16), (100, 300, 7), (1000, 5000, 166), (777, 9000, 165),
                                   (5000, 20000, 64), (300, 2000, 192), (130, 700, 100)])
def test_layer_fwd_bwd_input_space(N, E, K):
    got, ref = _fwd_bwd_case(N, E, K, algo=_abi.GEMM_INPUT, need_dx=False)
    assert got["dx"] is None
    _assert_close(got, ref, keys=KEYS)


def test_rows_longer_than_one_chunk_and_hub_rows_input_space():
    ei = synth.fraud_ring_skew(num_nodes=3000, background_edges=20000, num_hubs=4, hub_degree=2500, num_rings=20,
                               ring_len=16, seed=7)
    extra = torch.stack([torch.randint(0, 3000, (700,)), torch.full((700,), 11)])
    mid = torch.stack([torch.randint(0, 3000, (100,)), torch.full((100,), 12)])
    hub_src = torch.stack([torch.full((1500,), 13), torch.randint(0, 3000, (1500,))])
    ei = torch.cat([ei, extra, mid, hub_src], 1)
    for _ in range(3):      # the Gd prefetch is timing dependent: repeat
        _assert_close(*_fwd_bwd_case(3000, 0, 166, ei=ei, algo=_abi.GEMM_INPUT, need_dx=False), keys=KEYS)


def test_degree_skew_config5_input_space():
    N = 50_000
    ei = synth.fraud_ring_skew(num_nodes=N, background_edges=200_000, num_hubs=2, hub_degree=131_072, num_rings=50,
                               ring_len=64, seed=7)
    _assert_close(*_fwd_bwd_case(N, 0, 166, ei=ei, d_scale=1.0 / N, algo=_abi.GEMM_INPUT, need_dx=False), keys=KEYS)


def test_blocked_backward_matches_unblocked():
    """The backward edge pass processes destination rows in blocks (bounded Gd buffer): any block count gives the same dz."""
    N, K = 20_000, 166
    x, ei, _ = synth.elliptic_synth(num_nodes=N, num_edges=23_000, num_feats=K, num_steps=7, seed=0)
    W, a_s, a_d, b = (t.cuda() for t in seeded_params(K, 8, 64, seed=1))
    asf, adf = a_s.view(-1).contiguous(), a_d.view(-1).contiguous()
    g = build_csr(ei.cuda(), N)
    xg = Fn.in_pad_x(x.cuda())
    prep = Fn._aligned_u8(Fn.in_sizes(N, K)[0], xg.device)
    xmax = torch.zeros(16, device="cuda")
    a_src, a_dst = Fn.in_logits(xg, W, asf, adf, prep, xmax)
    Fn.in_prepare(W, K, xmax, prep)
    zimg, att = Fn.in_fwd(g, xg, a_src, a_dst, 0.2, prep)
    d_out = torch.randn(N, 64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)) / N
    ref = Fn.in_bwd_edges(g, xg, att, d_out, prep, 0.2, n_blocks=1)
    for nb in (2, 5):
        got = Fn.in_bwd_edges(g, xg, att, d_out, prep, 0.2, n_blocks=nb)
        assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1])


def test_attention_dropout_with_injected_mask_input_space():
    N, E, K, H, C = 400, 3000, 20, 8, 64
    W, a_s, a_d, b = seeded_params(K, H, C, seed=5)
    x = torch.randn(N, K, generator=torch.Generator().manual_seed(0))
    ei = synth.random_graph(N, E, seed=1)
    Ep = O.rewrite_self_loops(ei, N).size(1)
    keep = torch.rand(Ep, H, generator=torch.Generator().manual_seed(2)) >= 0.2
    conv = _layer(K, H, C, False, W, a_s, a_d, b, dropout=0.2, gemm_algo=_abi.GEMM_INPUT)
    out = conv(x.cuda(), ei.cuda(), dropout_mask=keep)
    leaves = [t.clone().requires_grad_(True) for t in (W, a_s, a_d, b)]
    ref, _ = O.gatconv_forward(x, ei, leaves[0], leaves[1], leaves[2], leaves[3], H, C, dropout_mask=keep, p=0.2)
    d_out = torch.randn(N, C, generator=torch.Generator().manual_seed(3)) / N
    out.backward(d_out.cuda()); ref.backward(d_out)
    assert maxabs(out, ref) <= 1e-5
    for g_, r_ in ((conv.lin_src.weight.grad, leaves[0].grad), (conv.att_src.grad, leaves[1].grad),
                   (conv.att_dst.grad, leaves[2].grad), (conv.bias.grad, leaves[3].grad)):
        assert maxabs(g_, r_) <= 1e-5


def test_elliptic_shape_full_size_layer1_input_space():
    """BASELINE config #2 shape, layer 1: N=203,769, E=234,355, K=166, against the fp64 oracle."""
    x, ei, _ = synth.elliptic_synth(seed=0)
    N, K = x.shape
    W, a_s, a_d, b = seeded_params(K, 8, 64, seed=1)
    conv = _layer(K, 8, 64, False, W, a_s, a_d, b, gemm_algo=_abi.GEMM_INPUT)
    out, (ei2, alpha) = conv(x.cuda(), ei.cuda(), return_attention_weights=True)
    ref_out, (ref_ei, ref_alpha) = O.gatconv_forward(x.double(), ei, W.double(), a_s.double(), a_d.double(), b.double(), 8, 64)
    assert torch.equal(ei2.cpu(), ref_ei)
    assert maxabs(out, ref_out) <= 1e-5 and maxabs(alpha, ref_alpha) <= 1e-5 and relerr(out, ref_out) <= 1e-5
    d_out = torch.randn(N, 64, generator=torch.Generator().manual_seed(3)) / N
    out.backward(d_out.cuda())
    cf = O.gatconv_backward_closed_form(x.double(), ei, W.double(), a_s.double(), a_d.double(), 8, 64, d_out.double(),
                                        need_dx=False)
    for got, ref in ((conv.lin_src.weight.grad, cf["dW"]), (conv.att_src.grad, cf["datt_src"]),
                     (conv.att_dst.grad, cf["datt_dst"]), (conv.bias.grad, cf["dbias"])):
        assert maxabs(got, ref) <= 1e-5 and relerr(got, ref) <= 1e-5


def test_input_space_rejects_what_it_cannot_do():
    conv = GATConv(16, 64, heads=8, concat=False, gemm_algo=_abi.GEMM_INPUT).cuda()
    ei = torch.zeros(2, 0, dtype=torch.long, device="cuda")
    with pytest.raises(_abi.GnnfdError):          # a gradient w.r.t. x is requested: hidden layers take the projected path
        conv(torch.randn(10, 16, device="cuda", requires_grad=True), ei)
    with pytest.raises(_abi.GnnfdError):
        GATConv(16, 64, heads=8, concat=True, gemm_algo=_abi.GEMM_INPUT).cuda()(torch.randn(10, 16, device="cuda"), ei)
    with pytest.raises(_abi.GnnfdError):          # unpadded rows handed straight to the C ABI
        x = torch.randn(10, 166, device="cuda")
        prep = Fn._aligned_u8(Fn.in_sizes(10, 166)[0], x.device)
        Fn.in_logits(x, torch.randn(512, 166, device="cuda"), torch.randn(512, device="cuda"), torch.randn(512, device="cuda"),
                     prep, torch.zeros(16, device="cuda"))


def test_fp16_pair_scale_is_range_safe():
    """Features far from unit scale (1e-6 .. 1e+6): the power-of-two scale keeps the fp16 pair inside its range."""
    for scale in (1e-6, 1e3, 1e6):
        N, E, K = 500, 3000, 166
        W, a_s, a_d, b = seeded_params(K, 8, 64, seed=2)
        x = torch.randn(N, K, generator=torch.Generator().manual_seed(0)) * scale
        Ws = W / scale                                   # keep the logits O(1)
        ei = synth.random_graph(N, E, seed=4)
        conv = _layer(K, 8, 64, False, Ws, a_s, a_d, b, gemm_algo=_abi.GEMM_INPUT)
        out = conv(x.cuda(), ei.cuda())
        ref, _ = O.gatconv_forward(x.double(), ei, Ws.double(), a_s.double(), a_d.double(), b.double(), 8, 64)
        assert relerr(out, ref) <= 1e-5, scale


# ---- in-kernel attention dropout (counter-based RNG) -----------------------------------------------------------------
@pytest.mark.parametrize("algo", [_abi.GEMM_SIMT, _abi.GEMM_INPUT])
def test_in_kernel_dropout_equals_its_exported_mask(algo):
    """Training-mode attention dropout draws its bits inside the kernels.  gnnfd_dropout_mask exports the same bits: a run
    with the RNG and a run with that mask injected must agree BIT FOR BIT in the output and in every gradient (so forward
    and backward see one mask), and the injected-mask path is the one checked against the oracle above."""
    N, E, K, H, C, p = 3000, 30000, 166, 8, 64, 0.2
    W, a_s, a_d, b = seeded_params(K, H, C, seed=5)
    x = torch.randn(N, K, generator=torch.Generator().manual_seed(0)).cuda()
    ei = torch.cat([synth.random_graph(N, E, seed=1), torch.stack([torch.randint(0, N, (900,)), torch.full((900,), 7)])], 1).cuda()
    conv = _layer(K, H, C, False, W, a_s, a_d, b, dropout=p, gemm_algo=algo).train()
    need_dx = algo != _abi.GEMM_INPUT
    d_out = torch.randn(N, C, generator=torch.Generator().manual_seed(3)).cuda() / N
    torch.manual_seed(123)
    xg = x.clone().requires_grad_(need_dx)
    out = conv(xg, ei)
    out.backward(d_out)
    seed = conv.last_dropout_seed
    grads = [p_.grad.clone() for p_ in conv.parameters()] + ([xg.grad.clone()] if need_dx else [])
    g = conv._graph(ei, N)
    keep, scale = Fn.dropout_mask(seed, p, g.n_edges, H, x.device)
    rate = float(keep.float().mean())
    assert abs(rate - (1 - p)) < 3e-3 and abs(scale * (1 - p) - 1) < 1e-4          # 264K draws: sigma = 8e-4
    per_head = keep.float().mean(0)
    assert float((per_head - (1 - p)).abs().max()) < 6e-3
    # heads and consecutive edges are uncorrelated
    kf = keep.float() - keep.float().mean()
    assert abs(float((kf[:, 0] * kf[:, 1]).mean())) < 3e-3 and abs(float((kf[1:, 0] * kf[:-1, 0]).mean())) < 3e-3
    conv.zero_grad()
    xg2 = x.clone().requires_grad_(need_dx)
    out2 = conv(xg2, ei, dropout_mask=keep)
    out2.backward(d_out)
    # same bits, same arithmetic; only the survivor scale differs: 65536/(65536 - thr) (the exact keep probability of the
    # 16-bit draw, 1.249995 for p = 0.2) against the mask path's 1/(1-p) = 1.25, i.e. 3.8e-6 relative
    assert relerr(out, out2) < 1e-5
    for ga, pb in zip(grads, list(conv.parameters()) + ([xg2] if need_dx else [])):
        assert relerr(ga, pb.grad) < 2e-5
    # a different seed gives a different mask; the same torch seed reproduces the run
    torch.manual_seed(123)
    assert torch.equal(conv(x, ei), out.detach())
    torch.manual_seed(124)
    assert not torch.equal(conv(x, ei), out.detach())
    k2, _ = Fn.dropout_mask(seed + 1, p, g.n_edges, H, x.device)
    assert 0.25 < float((k2 != keep).float().mean()) < 0.40                        # 2 p (1-p) = 0.32
