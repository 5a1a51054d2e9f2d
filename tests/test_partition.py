"""CPU tests of the multi-GPU host logic: LPT time-step sharding and the destination-range plan, including a
world_size-2 gloo run that executes the partitioned layer with the oracle's arithmetic and checks it against the
un-partitioned result (the CUDA kernels themselves are covered by the -m gpu tests)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gnn_fraud_detection_b200 import synth
from gnn_fraud_detection_b200.partition import DstRangePlan, lpt_assign, snapshot_batches
from oracle import pyg_gatconv as O


def test_lpt_assign_balances_and_covers():
    w = [5, 9, 1, 7, 3, 8, 2, 6, 4, 10]
    for G in (1, 2, 4, 8):
        bins = lpt_assign(w, G)
        assert sorted(i for b in bins for i in b) == list(range(len(w)))
        loads = [sum(w[i] for i in b) for b in bins]
        assert max(loads) - min(loads) <= max(w)


def test_snapshot_sharding_needs_no_communication():
    x, ei, ts = synth.elliptic_synth(num_nodes=3000, num_edges=3500, num_feats=5, num_steps=9, seed=2)
    seen_n, seen_e = 0, 0
    for r in range(4):
        xl, el, ids = snapshot_batches(x, ei, ts, r, 4)
        assert torch.equal(xl, x[ids])
        assert el.numel() == 0 or int(el.max()) < ids.numel()
        keep = torch.isin(ei[1], ids)
        assert torch.equal(ids[el], ei[:, keep])           # every edge of an owned step, order preserved, relabelled
        assert torch.all(torch.isin(ei[0, keep], ids))     # ... and both endpoints are local: no halo
        seen_n += ids.numel(); seen_e += el.size(1)
    assert seen_n == 3000 and seen_e == 3500


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_dst_range_plan_is_a_partition(world):
    N, E = 5000, 40000
    ei = synth.powerlaw_graph(N, E, seed=9)
    plan = DstRangePlan.build(ei, N, world)
    assert plan.start[0] == 0 and plan.start[-1] == N and torch.all(plan.start[1:] >= plan.start[:-1])
    pos = plan.to_pos(torch.arange(N))
    assert pos.unique().numel() == N and int(pos.max()) < world * plan.rows_padded
    ei2 = O.rewrite_self_loops(ei, N)
    per_rank = [plan.local_edges(ei, r) for r in range(world)]
    assert sum(e.size(1) for e in per_rank) == ei2.size(1)
    counts = torch.tensor([e.size(1) for e in per_rank], dtype=torch.float64)
    assert counts.max() <= ei2.size(1) / world + torch.bincount(ei2[1]).max()      # edge-balanced up to one row
    for r, e in enumerate(per_rank):
        lo, hi = int(plan.start[r]), int(plan.start[r + 1])
        m = (ei2[1] >= lo) & (ei2[1] < hi)
        assert torch.equal(e, torch.stack([plan.to_pos(ei2[0, m]), plan.to_pos(ei2[1, m])]))   # same order as edge_index'


def _worker(rank, world, port, N, E, K, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        H, C = 8, 64
        ei = synth.powerlaw_graph(N, E, seed=4)
        g = torch.Generator().manual_seed(0)
        x = torch.randn(N, K, generator=g, dtype=torch.float64)
        W = O.glorot_(torch.empty(H * C, K, dtype=torch.float64), g)
        a_s = O.glorot_(torch.empty(1, H, C, dtype=torch.float64), g)
        a_d = O.glorot_(torch.empty(1, H, C, dtype=torch.float64), g)
        b = torch.zeros(C, dtype=torch.float64)
        d_out = torch.randn(N, C, generator=g, dtype=torch.float64) / N
        plan = DstRangePlan.build(ei, N, world)
        P, lo_g, hi_g = plan.rows_padded, int(plan.start[rank]), int(plan.start[rank + 1])
        n_loc, n_pos = hi_g - lo_g, world * P
        le = plan.local_edges(ei, rank)                       # padded positions
        # forward: project own rows, all-gather into the padded buffer (what the NCCL path does)
        xw_own = torch.zeros(P, H * C, dtype=torch.float64)
        xw_own[:n_loc] = x[lo_g:hi_g] @ W.t()
        bufs = [torch.zeros_like(xw_own) for _ in range(world)]
        dist.all_gather(bufs, xw_own)
        xw_pos = torch.cat(bufs).view(n_pos, H, C)
        a_src = (xw_pos * a_s).sum(-1)
        a_dst = (xw_pos * a_d).sum(-1)
        src, dst = le[0], le[1]
        e = torch.nn.functional.leaky_relu(a_src[src] + a_dst[dst], 0.2)
        m = torch.full((n_pos, H), -float("inf"), dtype=torch.float64).scatter_reduce(0, dst[:, None].expand_as(e), e, "amax")
        m = torch.where(torch.isinf(m), torch.zeros_like(m), m)
        p = (e - m[dst]).exp()
        s = torch.zeros(n_pos, H, dtype=torch.float64).index_add_(0, dst, p) + 1e-16
        alpha = p / s[dst]
        out_pos = torch.zeros(n_pos, H, C, dtype=torch.float64).index_add_(0, dst, alpha[:, :, None] * xw_pos[src]).mean(1)
        out_loc = out_pos[rank * P: rank * P + n_loc] + b
        # backward: partial dxw over every source position, reduce-scatter (sum then take own chunk)
        dO = torch.zeros(n_pos, C, dtype=torch.float64)
        dO[rank * P: rank * P + n_loc] = d_out[lo_g:hi_g]
        dO_h = (dO / H)[:, None, :].expand(n_pos, H, C)
        d_alpha = (dO_h[dst] * xw_pos[src]).sum(-1)
        t = torch.zeros(n_pos, H, dtype=torch.float64).index_add_(0, dst, alpha * d_alpha)
        z = a_src[src] + a_dst[dst]
        dz = alpha * (d_alpha - t[dst]) * torch.where(z > 0, 1.0, 0.2)
        da_src = torch.zeros(n_pos, H, dtype=torch.float64).index_add_(0, src, dz)
        da_dst = torch.zeros(n_pos, H, dtype=torch.float64).index_add_(0, dst, dz)
        dxw = torch.zeros(n_pos, H, C, dtype=torch.float64).index_add_(0, src, alpha[:, :, None] * dO_h[dst])
        dxw = dxw + da_src[:, :, None] * a_s + da_dst[:, :, None] * a_d
        dist.all_reduce(dxw)                                   # reduce-scatter == all-reduce + own chunk
        dxw_loc = dxw[rank * P: rank * P + n_loc].reshape(n_loc, H * C)
        dW = dxw_loc.t() @ x[lo_g:hi_g]
        dist.all_reduce(dW)
        ref = O.gatconv_backward_closed_form(x, ei, W, a_s, a_d, H, C, d_out)
        ref_out, _ = O.gatconv_forward(x, ei, W, a_s, a_d, b, H, C)
        ret[rank] = (float((out_loc - ref_out[lo_g:hi_g]).abs().max()), float((dW - ref["dW"]).abs().max()),
                     float((dxw_loc - ref["dxw"][lo_g:hi_g]).abs().max()))
    finally:
        dist.destroy_process_group()


def test_partitioned_layer_equals_full_layer_world2_gloo():
    world = 2
    ret = mp.Manager().dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, 600, 5000, 9, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        e_out, e_dw, e_dxw = ret[r]
        assert e_out < 1e-9 and e_dw < 1e-9 and e_dxw < 1e-9, (r, ret[r])
