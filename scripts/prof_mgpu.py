"""Per-phase timing of the replicated-input multi-GPU layer (sync + barrier around every phase, so the sum is an
upper bound of the pipelined step).  torchrun --nproc-per-node G scripts/prof_mgpu.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from gnn_fraud_detection_b200 import GATConv, functional as Fn, synth, partition

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
acc = {}
def T(name, f):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize()
    acc.setdefault(name, []).append((time.perf_counter() - t0) * 1e3); return r
for it in range(3):
    xw, asf, adf = T("project_fwd(all rows)", lambda: Fn.project_fwd(x, W, a_s, a_d, H, C))
    a_dst = adf[lo:lo + P]
    out, rmax, rsum = T("gat_fwd", lambda: Fn.gat_fwd(g, xw, asf, a_dst, bias, H, C, 0.2, False))
    def ag():
        dO_pad = torch.zeros(P, C, device=dev); dO_pad[:n] = d_out
        dO_full = torch.empty(part.n_pos, C, device=dev); dist.all_gather_into_tensor(dO_full, dO_pad); return dO_full
    dO_full = T("all_gather dOut", ag)
    au, dz, dad = T("bwd_dst", lambda: Fn.gat_bwd_dst(g, xw, asf, a_dst, rmax, rsum, d_out, H, C, 0.2, False))
    def a2a():
        r_eg = torch.empty(part.n_recv, 2 * H, device=dev); dist.all_to_all_single(r_eg, au._base, part.recv_splits, part.send_splits); return r_eg
    r_eg = T("all_to_all edge grads", a2a)
    dpad = torch.zeros(P, H, device=dev); dpad[:n] = dad
    dxw, das = T("bwd_src(own sources)", lambda: Fn.gat_bwd_src(rg, r_eg[:, :H], r_eg[:, H:], dO_full, a_s, a_d, dpad, H, C, False))
    grads = T("project_bwd(own rows)", lambda: Fn.project_bwd(x[lo:lo + n], W, dxw[:n], xw[lo:lo + n], das[:n], dad, d_out, H, C, C, False))
    T("all_reduce grads", lambda: dist.all_reduce(torch.cat([grads[0].reshape(-1), grads[1], grads[2], grads[3]])))
if rank == 0:
    tot = 0.0
    for k, v in acc.items():
        print(f"{k:28s} {v[-1]:8.2f} ms"); tot += v[-1]
    print(f"{'sum':28s} {tot:8.2f} ms   (world={world}, n_recv={part.n_recv}, send_splits={part.send_splits[:4]}...)")
dist.destroy_process_group()
