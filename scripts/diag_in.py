"""Stage-by-stage check of the input-space first-layer path against the fp64 oracle (run on the GPU box).

  python scripts/diag_in.py            # a few small graphs, every intermediate compared
Prints max-abs / relative-L2 errors per stage so that one GPU call localises a wrong kernel.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from gnn_fraud_detection_b200 import _abi, build_csr, functional as Fn, synth  # noqa: E402
from oracle import pyg_gatconv as O  # noqa: E402
from util import seeded_params  # noqa: E402

H, C = 8, 64


def decode_zimg(zimg: torch.Tensor, n: int, K: int, sx: float) -> torch.Tensor:
    """fp16-pair image -> Z [n, H, KP] float64 (csrc/in_common.cuh layout)."""
    KP = (K + 7) & ~7
    NKB = KP // 8
    raw = zimg.cpu().numpy().view(np.float16)
    n_tiles = (n + 127) // 128
    r = np.arange(128)[:, None]
    e = np.arange(64)[None, :]
    off = ((r >> 3) * 1024 + (r & 7) * 128 + (((e >> 3) ^ (r & 7)) << 4) + (e & 7) * 2) // 2      # fp16 index inside a plane
    Z = np.zeros((n_tiles * 128, NKB * 64), dtype=np.float64)
    for t in range(n_tiles):
        for kb in range(NKB):
            base = ((t * NKB + kb) * 2) * 8192
            hi = raw[base: base + 8192][off].astype(np.float64)
            lo = raw[base + 8192: base + 16384][off].astype(np.float64)
            Z[t * 128:(t + 1) * 128, kb * 64:(kb + 1) * 64] = hi + lo
    return torch.from_numpy(Z[:n] / sx).view(n, H, KP)


def err(name, got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    ma = float((got - ref).abs().max()) if got.numel() else 0.0
    rel = float((got - ref).norm() / (ref.norm() + 1e-300))
    flag = "" if (rel < 1e-5 or ma < 1e-7) else "   <<<<<< BAD"
    print(f"  {name:10s} max-abs {ma:.3e}  rel-L2 {rel:.3e}  (|ref|max {float(ref.abs().max()) if ref.numel() else 0:.3e}){flag}")
    return rel


def case(name, N, K, ei, seed=0, n_blocks=None, keep=None, p=0.0):
    print(f"== {name}: N={N} K={K} E={ei.size(1)} blocks={n_blocks} dropout={p}")
    dev = torch.device("cuda")
    W, a_s, a_d, b = seeded_params(K, H, C, False, seed=seed + 1)
    x = torch.randn(N, K, generator=torch.Generator().manual_seed(seed))
    d_out = torch.randn(N, C, generator=torch.Generator().manual_seed(seed + 3)) / max(N, 1)
    # ---- oracle, fp64
    xd, Wd, asd, add_ = x.double(), W.double(), a_s.double(), a_d.double()
    out_ref, (ei2, alpha) = O.gatconv_forward(xd, ei, Wd, asd, add_, b.double(), H, C, False, dropout_mask=keep, p=p)
    src, dst = ei2[0], ei2[1]
    xw = (xd @ Wd.t()).view(N, H, C)
    a_src_ref = (xw * asd.view(1, H, C)).sum(-1)
    a_dst_ref = (xw * add_.view(1, H, C)).sum(-1)
    au = alpha if keep is None else alpha * keep.double() / (1 - p)
    Z_ref = torch.zeros(N, H, K, dtype=torch.float64).index_add_(0, dst, au[:, :, None] * xd[src][:, None, :])
    leaves = [t.clone().requires_grad_(True) for t in (Wd, asd, add_, b.double())]
    o2, _ = O.gatconv_forward(xd, ei, leaves[0], leaves[1], leaves[2], leaves[3], H, C, False, dropout_mask=keep, p=p)
    o2.backward(d_out.double())
    # ---- device
    g = build_csr(ei.to(dev), N)
    xg, Wg = Fn.in_pad_x(x.to(dev)), W.to(dev)
    asg, adg, bg = a_s.to(dev).view(-1).contiguous(), a_d.to(dev).view(-1).contiguous(), b.to(dev)
    keep_g = None if keep is None else keep.to(dev).to(torch.uint8).contiguous()
    pb, zb, F, _ = Fn.in_sizes(N, K)
    prep = Fn._aligned_u8(pb, dev)
    xmax = torch.zeros(16, device=dev)
    a_src, a_dst = Fn.in_logits(xg, Wg, asg, adg, prep, xmax)
    err("a_src", a_src, a_src_ref); err("a_dst", a_dst, a_dst_ref)
    print("  xmax", float(xmax[0]), "vs", float(x.abs().max()) if N else 0.0)
    Fn.in_prepare(Wg, K, xmax, prep)
    scal = prep[:16].view(torch.float32).cpu()
    print("  scal", scal.tolist())
    zimg, att = Fn.in_fwd(g, xg, a_src, a_dst, 0.2, prep, keep_g, p)
    rowmax, rowsum = att.rowmax, att.rowsum
    torch.cuda.synchronize()
    Z = decode_zimg(zimg, N, K, float(scal[0]))
    err("Z", Z[:, :, :K], Z_ref)
    out = Fn.in_out(zimg, N, K, prep, bg)
    torch.cuda.synchronize()
    err("out", out, out_ref)
    dog = d_out.to(dev)
    gd = torch.empty(N, F, device=dev)
    import ctypes
    gb = ctypes.c_size_t()
    _abi.check(_abi.lib().gnnfd_in_bwd_gd_workspace_bytes(N, ctypes.byref(gb)))
    gws = Fn._aligned_u8(gb.value, dev)
    _abi.check(_abi.lib().gnnfd_in_bwd_gd(dog.data_ptr(), N, K, prep.data_ptr(), gd.data_ptr(), gws.data_ptr(), gws.numel(),
                                          torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    KP = F // H
    Gd_ref = torch.einsum("nc,hck->nhk", d_out.double() / H, Wd.view(H, C, K))
    err("Gd", gd.view(N, H, KP)[:, :, :K], Gd_ref)
    if KP > K:
        print("  Gd pad max", float(gd.view(N, H, KP)[:, :, K:].abs().max()))
    dz, da_dst = Fn.in_bwd_edges(g, xg, att, dog, prep, 0.2, keep_g, p, n_blocks=n_blocks)
    torch.cuda.synchronize()
    # closed form pieces in fp64 (with dropout: d_alpha path scaled by keep/(1-p))
    dO_h = (d_out.double() / H)[:, None, :].expand(N, H, C)
    d_alpha = (dO_h[dst] * xw[src]).sum(-1)
    if keep is not None:
        d_alpha = d_alpha * keep.double() / (1 - p)
    z = a_src_ref[src] + a_dst_ref[dst]
    t = torch.zeros(N, H, dtype=torch.float64).index_add_(0, dst, alpha * d_alpha)
    dz_ref = alpha * (d_alpha - t[dst]) * torch.where(z > 0, 1.0, 0.2)
    perm, c2c = g.perm.long().cpu(), g.csr2csc.long().cpu()
    dz_csr = dz.cpu()[c2c]                 # CSR position e -> source-major slot
    err("dz", dz_csr, dz_ref[perm])
    err("da_dst", da_dst, torch.zeros(N, H, dtype=torch.float64).index_add_(0, dst, dz_ref))
    da_src = Fn.in_dasrc(g, dz)
    err("da_src", da_src, torch.zeros(N, H, dtype=torch.float64).index_add_(0, src, dz_ref))
    dW, datt_s, datt_d, dbias = Fn.in_bwd_params(zimg, dog, xg, Wg, asg, adg, da_src, da_dst, prep)
    torch.cuda.synchronize()
    r = [err("dW", dW, leaves[0].grad), err("datt_src", datt_s, leaves[1].grad.view(-1)),
         err("datt_dst", datt_d, leaves[2].grad.view(-1)), err("dbias", dbias, leaves[3].grad)]
    return max(r)


def big_case(N=200_000, E=2_000_000, K=166):
    """dW / datt accuracy at a size where the node-reduction chains are long: fp64 reference evaluated on the GPU."""
    print(f"== big: N={N} E={E}")
    dev = torch.device("cuda")
    W, a_s, a_d, b = seeded_params(K, H, C, False, seed=1)
    ei = synth.powerlaw_graph(N, E, seed=5, device=dev)
    x = torch.randn(N, K, device=dev, generator=torch.Generator(device=dev).manual_seed(0))
    d_out = torch.randn(N, C, device=dev, generator=torch.Generator(device=dev).manual_seed(2)) / N
    with torch.device("cuda"):
        cf = O.gatconv_backward_closed_form(x.double(), ei, W.double().cuda(), a_s.double().cuda(), a_d.double().cuda(), H, C,
                                            d_out.double(), False, need_dx=False)
        ro, _ = O.gatconv_forward(x.double(), ei, W.double().cuda(), a_s.double().cuda(), a_d.double().cuda(), b.double().cuda(), H, C, False)
    g = build_csr(ei, N)
    xg, Wg = Fn.in_pad_x(x), W.to(dev)
    asg, adg, bg = a_s.to(dev).view(-1).contiguous(), a_d.to(dev).view(-1).contiguous(), b.to(dev)
    prep = Fn._aligned_u8(Fn.in_sizes(N, K)[0], dev)
    xmax = torch.zeros(16, device=dev)
    a_src, a_dst = Fn.in_logits(xg, Wg, asg, adg, prep, xmax)
    Fn.in_prepare(Wg, K, xmax, prep)
    zimg, att = Fn.in_fwd(g, xg, a_src, a_dst, 0.2, prep)
    out = Fn.in_out(zimg, N, K, prep, bg)
    err("out", out, ro)
    dz, da_dst = Fn.in_bwd_edges(g, xg, att, d_out, prep, 0.2)
    da_src = Fn.in_dasrc(g, dz)
    err("da_src", da_src, cf["da_src"]); err("da_dst", da_dst, cf["da_dst"])
    dW, ds, dd, db = Fn.in_bwd_params(zimg, d_out, xg, Wg, asg, adg, da_src, da_dst, prep)
    err("dW", dW, cf["dW"]); err("datt_src", ds, cf["datt_src"].view(-1)); err("datt_dst", dd, cf["datt_dst"].view(-1)); err("dbias", db, cf["dbias"])
    # the projected-feature path on the same inputs
    xw, a_src2, a_dst2 = Fn.project_fwd(x, Wg, asg, adg, H, C)
    out2, rm2, rs2 = Fn.gat_fwd(g, xw, a_src2, a_dst2, bg, H, C, 0.2, False)
    dxw, das2, dad2 = Fn.gat_bwd(g, xw, a_src2, a_dst2, rm2, rs2, d_out, asg, adg, H, C, 0.2, False)
    dW2, ds2, dd2, db2, _ = Fn.project_bwd(x, Wg, dxw, xw, das2, dad2, d_out, H, C, C, False)
    print("  projected-feature path:")
    err("out", out2, ro); err("dW", dW2, cf["dW"]); err("datt_src", ds2, cf["datt_src"].view(-1)); err("dbias", db2, cf["dbias"])


def main():
    torch.manual_seed(0)
    if os.environ.get("DIAG_BIG"):
        big_case()
        return
    worst = 0.0
    worst = max(worst, case("tiny", 3, 5, torch.tensor([[0, 1, 2, 2], [1, 2, 0, 2]])))
    worst = max(worst, case("small", 300, 166, synth.random_graph(300, 1500, seed=2)))
    worst = max(worst, case("odd K", 777, 165, synth.random_graph(777, 9000, seed=3)))
    if os.environ.get("DIAG_QUICK"):
        print("WORST (quick):", worst)
        return
    worst = max(worst, case("K=64", 1000, 64, synth.random_graph(1000, 4000, seed=4)))
    ei = synth.fraud_ring_skew(num_nodes=3000, background_edges=20000, num_hubs=4, hub_degree=2500, num_rings=20, ring_len=16, seed=7)
    extra = torch.stack([torch.randint(0, 3000, (700,)), torch.full((700,), 11)])
    mid = torch.stack([torch.randint(0, 3000, (100,)), torch.full((100,), 12)])
    ei = torch.cat([ei, extra, mid], 1)
    worst = max(worst, case("hubs", 3000, 166, ei))
    worst = max(worst, case("hubs blocked x3", 3000, 166, ei, n_blocks=3))
    ei = synth.random_graph(2000, 16000, seed=5)
    Ep = O.rewrite_self_loops(ei, 2000).size(1)
    keep = torch.rand(Ep, H, generator=torch.Generator().manual_seed(2)) >= 0.2
    worst = max(worst, case("dropout", 2000, 166, ei, keep=keep, p=0.2))
    x, ei, _ = synth.elliptic_synth(num_nodes=20000, num_edges=23000, num_feats=166, num_steps=7, seed=0)
    worst = max(worst, case("elliptic-ish", 20000, 166, ei, n_blocks=2))
    print("WORST relative error over final gradients:", worst)


if __name__ == "__main__":
    main()
