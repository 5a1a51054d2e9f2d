#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/prof2.py <<'PY'
import os, sys, torch, torch.distributed as dist, time
sys.path.insert(0, os.getcwd())
from gnn_fraud_detection_b200 import GATConv, _abi, functional as Fn, synth, partition
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
N, E, K, H, C = 20_000_000, 200_000_000, 166, 8, 64
ei = synth.powerlaw_graph(N, E, seed=1234, device=dev)
part = partition.ReplicatedInputPartition.build(ei, N, rank, world, dev); del ei
x = torch.randn(part.n_pos, K, device=dev, generator=torch.Generator(device=dev).manual_seed(0))
torch.manual_seed(1); conv = GATConv(K, C, heads=H, concat=False).to(dev)
W = conv.lin_src.weight.detach(); a_s = conv.att_src.detach().view(-1).contiguous(); a_d = conv.att_dst.detach().view(-1).contiguous(); bias = conv.bias.detach()
d_out = torch.full((part.n_local, C), 1.0 / N, device=dev)
g, rg, P, n, lo = part.graph, part.rgraph, part.rows_padded, part.n_local, rank * part.rows_padded
def T(name, f):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    print(f"[rank {rank}] {name}: {dt:.2f} ms", flush=True); return r
deg = (rg.colptr[1:] - rg.colptr[:-1])
print(f"[rank {rank}] n_recv={part.n_recv} P={P} n_local={n} maxdeg={int(deg.max())} hubs={rg.c.hub_src.n_hub} items={rg.c.items_src.n_items} target={rg.c.items_src.target} colptr_last={int(rg.colptr[-1])} recv_splits={part.recv_splits} send_splits={part.send_splits}", flush=True)
ldeg = (g.colptr[1:] - g.colptr[:-1])
print(f"[rank {rank}] local CSC: maxdeg={int(ldeg.max())} hubs={g.c.hub_src.n_hub} items={g.c.items_src.n_items}", flush=True)
for it in range(2):
    xw, asf, adf = T("project_fwd", lambda: Fn.project_fwd(x, W, a_s, a_d, H, C))
    a_dst = adf[lo:lo + P]
    out, rmax, rsum = T("gat_fwd", lambda: Fn.gat_fwd(g, xw, asf, a_dst, bias, H, C, 0.2, False))
    def ag():
        dO_pad = torch.zeros(P, C, device=dev); dO_pad[:n] = d_out
        dO_full = torch.empty(part.n_pos, C, device=dev); dist.all_gather_into_tensor(dO_full, dO_pad); return dO_full
    dO_full = T("allgather_dOut", ag)
    au, dz, dad = T("bwd_dst", lambda: Fn.gat_bwd_dst(g, xw, asf, a_dst, rmax, rsum, d_out, H, C, 0.2, False))
    def a2a():
        ra = torch.empty(part.n_recv, H, device=dev); rz = torch.empty(part.n_recv, H, device=dev)
        dist.all_to_all_single(ra, au, part.recv_splits, part.send_splits); dist.all_to_all_single(rz, dz, part.recv_splits, part.send_splits); return ra, rz
    ra, rz = T("all_to_all", a2a)
    dpad = torch.zeros(P, H, device=dev); dpad[:n] = dad
    dxw, das = T("bwd_src", lambda: Fn.gat_bwd_src(rg, ra, rz, dO_full, a_s, a_d, dpad, H, C, False))
    dpos = torch.zeros(part.n_pos, H, device=dev)
    _ = T("bwd_src_localgraph", lambda: Fn.gat_bwd_src(g, au, dz, d_out, a_s, a_d, dpos, H, C, False)); del _
    grads = T("project_bwd", lambda: Fn.project_bwd(x[lo:lo + n], W, dxw[:n], xw[lo:lo + n], das[:n], dad, d_out, H, C, C, False))
    T("all_reduce", lambda: dist.all_reduce(torch.cat([grads[0].reshape(-1), grads[1], grads[2], grads[3]])))
dist.destroy_process_group()
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 /tmp/prof2.py > gpurun_out/prof2.log 2>&1; grep "rank" gpurun_out/prof2.log | tail -20
