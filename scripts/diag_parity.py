"""Diagnostic: relative errors of every intermediate of one layer, and of a 2-layer GAT training step,
against the fp64 oracle (the fp32 oracle's own error is printed beside ours)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from gnn_fraud_detection_b200 import GAT, _abi, build_csr, functional as Fn, synth
from oracle import pyg_gatconv as O
from util import seeded_params

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-300)), float((a - b).abs().max())

def layer(N, E, K, seed=0):
    H, C = 8, 64
    W, a_s, a_d, b = seeded_params(K, H, C, False, seed=seed + 1)
    x = torch.randn(N, K, generator=torch.Generator().manual_seed(seed))
    ei = synth.elliptic_synth(N, E, 1, seed=seed)[1] if N > 5000 else synth.random_graph(N, E, seed=seed + 2)
    d_out = torch.randn(N, C, generator=torch.Generator().manual_seed(seed + 3)) / N
    cf = O.gatconv_backward_closed_form(x.double(), ei, W.double(), a_s.double(), a_d.double(), H, C, d_out.double())
    cf32 = O.gatconv_backward_closed_form(x, ei, W, a_s, a_d, H, C, d_out)
    g = build_csr(ei.cuda(), N)
    xg, Wg, asg, adg, bg, dg = x.cuda(), W.cuda(), a_s.cuda().view(-1), a_d.cuda().view(-1), b.cuda(), d_out.cuda()
    xw, a_src, a_dst = Fn.project_fwd(xg, Wg, asg, adg, H, C, torch.float32, _abi.GEMM_SIMT)
    out, rowmax, rowsum = Fn.gat_fwd(g, xw, a_src, a_dst, bg, H, C, 0.2, False)
    alpha = Fn.gat_alpha(g, a_src, a_dst, rowmax, rowsum, H, 0.2)
    dxw, da_src, da_dst = Fn.gat_bwd(g, xw, a_src, a_dst, rowmax, rowsum, dg, asg, adg, H, C, 0.2, False)
    dW, datt_s, datt_d, dbias, dx = Fn.project_bwd(xg, Wg, dxw, xw, da_src, da_dst, dg, H, C, C, True, _abi.GEMM_SIMT)
    perm = g.perm.long().cpu()
    print(f"--- layer N={N} E={E} K={K}: (rel-L2, max-abs) ours vs fp64 | fp32-oracle vs fp64")
    def row(name, ours, ref64, ref32):
        print(f"{name:10s} ours {rel(ours, ref64)}   oracle32 {rel(ref32, ref64)}")
    xw64 = (x.double() @ W.double().t())
    row("xw", xw, xw64, x @ W.t())
    row("alpha", alpha.cpu(), cf["alpha"][perm], cf32["alpha"][perm])
    row("da_src", da_src, cf["da_src"], cf32["da_src"])
    row("da_dst", da_dst, cf["da_dst"], cf32["da_dst"])
    row("dxw", dxw, cf["dxw"], cf32["dxw"])
    row("dW", dW, cf["dW"], cf32["dW"])
    row("datt_src", datt_s.view(1, H, C), cf["datt_src"], cf32["datt_src"])
    row("datt_dst", datt_d.view(1, H, C), cf["datt_dst"], cf32["datt_dst"])
    row("dbias", dbias, cf["dbias"], cf32["dbias"])
    row("dx", dx, cf["dx"], cf32["dx"])

def model():
    x, ei, _ = synth.elliptic_synth(num_nodes=20_000, num_edges=23_000, num_feats=166, seed=0)
    torch.manual_seed(0)
    ref = O.OracleGAT(166, 64, 1, num_layers=2, dropout=0.0)
    ref64 = O.OracleGAT(166, 64, 1, num_layers=2, dropout=0.0).double()
    ref64.load_state_dict(ref.state_dict())
    ours = GAT(166, 64, 1, num_layers=2, dropout=0.0)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ours = ours.cuda()
    y = (torch.rand(x.size(0), 1, generator=torch.Generator().manual_seed(1)) < 0.1).float()
    crit = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor(50.0))
    crit(ref(x, ei), y).backward()
    torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor(50.0).double())(ref64(x.double(), ei), y.double()).backward()
    crit.cuda()(ours(x.cuda(), ei.cuda()), y.cuda()).backward()
    print("--- 2-layer GAT training step: (rel-L2, max-abs) ours vs fp64 | fp32-oracle vs fp64")
    for (n, p64), (_, p32), (_, po) in zip(ref64.named_parameters(), ref.named_parameters(), ours.named_parameters()):
        print(f"{n:32s} ours {rel(po.grad, p64.grad)}   oracle32 {rel(p32.grad, p64.grad)}")

layer(1000, 5000, 166)
layer(20000, 23000, 166)
layer(20000, 23000, 64)
model()
