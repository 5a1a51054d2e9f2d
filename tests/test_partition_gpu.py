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
        from gnn_fraud_detection_b200.partition import DstRangePartition, ReplicatedInputPartition, snapshot_batches
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
        rel_dw = float((dW2 - dW).norm() / dW.norm())
        # variant (c): replicated input, per-edge gradient exchange
        part2 = ReplicatedInputPartition.build(ei, N, rank, world, dev)
        x_pos = torch.zeros(part2.n_pos, K, device=dev)
        x_pos[part2.plan.to_pos(torch.arange(N, device=dev))] = x
        o3, (dW3, ds3, dd3, db3) = part2.layer_fwd_bwd(x_pos, W, a_s, a_d, bias, d_out[lo:hi].contiguous(), H, C,
                                                      torch.float32, _abi.GEMM_AUTO)
        errs += [float((o3 - out[lo:hi]).abs().max()), float((dW3 - dW).abs().max()), float((ds3 - datt_s).abs().max()),
                 float((dd3 - datt_d).abs().max()), float((db3 - dbias).abs().max())]
        rel_dw = max(rel_dw, float((dW3 - dW).norm() / dW.norm()))
        # time-step sharding: a rank's block-diagonal batch reproduces the full-graph rows it owns
        xs, es, ts = synth.elliptic_synth(num_nodes=40_000, num_edges=46_000, num_feats=K, seed=0, device=dev)
        full = conv.eval()(xs, es)
        xl, el, ids = snapshot_batches(xs, es, ts, rank, world)
        loc = conv(xl.contiguous(), el)
        errs.append(float((loc - full[ids]).abs().max()))
        ret[rank] = (errs, rel_dw)
    finally:
        dist.destroy_process_group()


def test_dst_range_partition_matches_single_gpu():
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, 29650 + os.getpid() % 1000, ret), nprocs=world, join=True)
    for r in range(world):
        errs, rel_dw = ret[r]
        assert max(errs) <= 1e-5 and rel_dw <= 1e-5, (r, errs, rel_dw)
