"""Multi-GPU parity (needs >= 2 GPUs; skipped on the single-GPU box): the destination-range partition with NCCL
all-gather / reduce-scatter must reproduce the single-GPU layer, and time-step sharding the full-graph output."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from gnn_fraud_detection_b200 import _abi, build_csr, functional as Fn, synth
        from gnn_fraud_detection_b200.partition import (DstRangePartition, InputSpacePartition, ReplicatedInputPartition,
                                                         snapshot_batches)
        from gnn_fraud_detection_b200 import GATConv
        H, C, K, N, E = 8, 64, 166, 200_000, 2_000_000
        ei = synth.powerlaw_graph(N, E, seed=5, device=dev)
        x = torch.randn(N, K, device=dev, generator=torch.Generator(device=dev).manual_seed(0))
        torch.manual_seed(1)
        conv = GATConv(K, C, heads=H, concat=False).to(dev)
        W, bias = conv.lin_src.weight.detach(), conv.bias.detach()
        a_s, a_d = conv.att_src.detach().view(-1).contiguous(), conv.att_dst.detach().view(-1).contiguous()
        d_out = torch.randn(N, C, device=dev, generator=torch.Generator(device=dev).manual_seed(2)) / N
        # single-GPU reference on every rank
        g = build_csr(ei, N)
        xw, a_src, a_dst = Fn.project_fwd(x, W, a_s, a_d, H, C)
        out, rowmax, rowsum = Fn.gat_fwd(g, xw, a_src, a_dst, bias, H, C, 0.2, False)
        dxw, da_src, da_dst = Fn.gat_bwd(g, xw, a_src, a_dst, rowmax, rowsum, d_out, a_s, a_d, H, C, 0.2, False)
        dW, datt_s, datt_d, dbias, _ = Fn.project_bwd(x, W, dxw, xw, da_src, da_dst, d_out, H, C, C, False)
        # partitioned
        part = DstRangePartition.build(ei, N, rank, world, dev)
        lo, hi = int(part.plan.start[rank]), int(part.plan.start[rank + 1])
        x_loc = torch.zeros(part.rows_padded, K, device=dev)
        x_loc[:hi - lo] = x[lo:hi]
        o2, (dW2, ds2, dd2, db2) = part.layer_fwd_bwd(x_loc, W, a_s, a_d, bias, d_out[lo:hi].contiguous(), H, C,
                                                     torch.float32, _abi.GEMM_AUTO)
        errs = [float((o2 - out[lo:hi]).abs().max()), float((dW2 - dW).abs().max()), float((ds2 - datt_s).abs().max()),
                float((dd2 - datt_d).abs().max()), float((db2 - dbias).abs().max())]
        rels = {"allgather dW vs 1 GPU": float((dW2 - dW).norm() / dW.norm())}
        # variant (c): replicated input, per-edge gradient exchange
        part2 = ReplicatedInputPartition.build(ei, N, rank, world, dev)
        x_pos = torch.zeros(part2.n_pos, K, device=dev)
        x_pos[part2.plan.to_pos(torch.arange(N, device=dev))] = x
        o3, (dW3, ds3, dd3, db3) = part2.layer_fwd_bwd(x_pos, W, a_s, a_d, bias, d_out[lo:hi].contiguous(), H, C,
                                                      torch.float32, _abi.GEMM_AUTO)
        errs += [float((o3 - out[lo:hi]).abs().max()), float((dW3 - dW).abs().max()), float((ds3 - datt_s).abs().max()),
                 float((dd3 - datt_d).abs().max()), float((db3 - dbias).abs().max())]
        rels["replicate dW vs 1 GPU"] = float((dW3 - dW).norm() / dW.norm())
        # input-space formulation (x replicated, only [N,H] logit vectors cross NVLink)
        part3 = InputSpacePartition.build(ei, N, rank, world, dev)
        x16 = torch.zeros(part3.n_pos, Fn.in_sizes(0, K)[3], device=dev)
        x16[part3.plan.to_pos(torch.arange(N, device=dev)), :K] = x
        o4, (dW4, ds4, dd4, db4) = part3.layer_fwd_bwd(x16[:, :K], W, a_s, a_d, bias, d_out[lo:hi].contiguous(), H, C)
        errs += [float((o4 - out[lo:hi]).abs().max()), float((dW4 - dW).abs().max()), float((ds4 - datt_s).abs().max()),
                 float((dd4 - datt_d).abs().max()), float((db4 - dbias).abs().max())]
        rels["input-space dW vs 1 GPU projected-feature"] = float((dW4 - dW).norm() / dW.norm())
        # the same layer with every exchange mechanism: NCCL collectives, per-peer NVLink stores / loads, NVSwitch multicast
        modes = {part3.exchange: (o4, dW4, ds4, dd4)}
        for mode in ("nccl", "peer", "multicast"):
            if mode in modes:
                continue
            os.environ["GNNFD_EXCHANGE"] = mode
            try:
                pm = InputSpacePartition.build(ei, N, rank, world, dev)
                om, (dWm, dsm, ddm, dbm) = pm.layer_fwd_bwd(x16[:, :K], W, a_s, a_d, bias, d_out[lo:hi].contiguous(), H, C)
                assert pm.exchange == mode
                modes[mode] = (om, dWm, dsm, ddm)
            except RuntimeError as ex:
                if mode != "multicast" or "multicast" not in str(ex):
                    raise
            finally:
                os.environ.pop("GNNFD_EXCHANGE", None)
        assert "nccl" in modes and "peer" in modes, list(modes)
        for mode, (om, dWm, dsm, ddm) in modes.items():
            errs += [float((om - out[lo:hi]).abs().max()), float((dWm - dW).abs().max()), float((dsm - datt_s).abs().max()),
                     float((ddm - datt_d).abs().max())]
            rels[f"input-space[{mode}] dW vs 1 GPU projected-feature"] = float((dWm - dW).norm() / dW.norm())
        rels["exchange modes covered: " + ",".join(sorted(modes))] = 0.0
        # a second step through the same buffers (stale data from step 1 must not leak into step 2)
        o4b, (dW4b, _, _, _) = part3.layer_fwd_bwd(x16[:, :K], W * 0.5, a_s, a_d, bias, d_out[lo:hi].contiguous(), H, C)
        o4c, (dW4c, _, _, _) = part3.layer_fwd_bwd(x16[:, :K], W, a_s, a_d, bias, d_out[lo:hi].contiguous(), H, C)
        errs += [float((o4c - o4).abs().max()), float((dW4c - dW4).abs().max())]
        # ... and directly against the fp64 CPU oracle on a graph it can hold
        from oracle import pyg_gatconv as O
        N2, E2 = 20_000, 200_000
        ei2 = synth.powerlaw_graph(N2, E2, seed=6, device=dev)
        x2 = torch.randn(N2, K, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
        d2 = torch.randn(N2, C, device=dev, generator=torch.Generator(device=dev).manual_seed(4)) / N2
        p5 = InputSpacePartition.build(ei2, N2, rank, world, dev)
        lo2, hi2 = int(p5.plan.start[rank]), int(p5.plan.start[rank + 1])
        x2p = torch.zeros(p5.n_pos, Fn.in_sizes(0, K)[3], device=dev)
        x2p[p5.plan.to_pos(torch.arange(N2, device=dev)), :K] = x2
        o5, (dW5, ds5, dd5, db5) = p5.layer_fwd_bwd(x2p[:, :K], W, a_s, a_d, bias, d2[lo2:hi2].contiguous(), H, C)
        torch.set_num_threads(4)
        Wc, asc, adc, bc = W.double().cpu(), conv.att_src.detach().double().cpu(), conv.att_dst.detach().double().cpu(), bias.double().cpu()
        ro, _ = O.gatconv_forward(x2.double().cpu(), ei2.cpu(), Wc, asc, adc, bc, H, C, False)
        cf = O.gatconv_backward_closed_form(x2.double().cpu(), ei2.cpu(), Wc, asc, adc, H, C, d2.double().cpu(), False, need_dx=False)
        errs += [float((o5.double().cpu() - ro[lo2:hi2]).abs().max()), float((dW5.double().cpu() - cf["dW"]).abs().max()),
                 float((ds5.double().cpu() - cf["datt_src"].view(-1)).abs().max()),
                 float((dd5.double().cpu() - cf["datt_dst"].view(-1)).abs().max()),
                 float((db5.double().cpu() - cf["dbias"]).abs().max())]
        rels["input-space dW vs fp64 oracle"] = float((dW5.double().cpu() - cf["dW"]).norm() / cf["dW"].norm())
        rels["input-space out vs fp64 oracle"] = float((o5.double().cpu() - ro[lo2:hi2]).norm() / ro[lo2:hi2].norm())
        # time-step sharding: a rank's block-diagonal batch reproduces the full-graph rows it owns
        xs, es, ts = synth.elliptic_synth(num_nodes=40_000, num_edges=46_000, num_feats=K, seed=0, device=dev)
        full = conv.eval()(xs, es)
        xl, el, ids = snapshot_batches(xs, es, ts, rank, world)
        loc = conv(xl.contiguous(), el)
        errs.append(float((loc - full[ids]).abs().max()))
        ret[rank] = (errs, rels)
    finally:
        dist.destroy_process_group()


def test_dst_range_partition_matches_single_gpu():
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, 29650 + os.getpid() % 1000, ret), nprocs=world, join=True)
    for r in range(world):
        errs, rels = ret[r]
        print(r, rels)
        assert max(errs) <= 1e-5, (r, errs)
        # two fp32 implementations with different summation orders are each ~1e-6 from the truth on a 200K-node reduction;
        # the comparisons against the fp64 oracle are the tight ones
        for name, v in rels.items():
            assert v <= (1e-5 if "oracle" in name else 3e-5), (r, name, v)
