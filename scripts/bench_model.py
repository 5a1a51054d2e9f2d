"""Model-level timings for BASELINE.json configs #1-#3 (not the headline metric): the reference's GAT / TemporalGNN
classes with the B200 layer vs the same classes over the CPU oracle layer (PyG formulation) on the host cores."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_fraud_detection_b200 import GAT, TemporalGNN, synth
from oracle import pyg_gatconv as O

def gpu_time(fn, warm=3, reps=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]

def graph_time(fn, warm=3, reps=20):
    """Same callable captured once into a CUDA graph (static inputs) and replayed: removes launch gaps."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warm): fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return gpu_time(g.replay, warm=3, reps=reps)

def cpu_time(fn, reps=2):
    fn(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e3

def main():
    torch.set_num_threads(os.cpu_count() or 1)
    x, ei, ts = synth.elliptic_synth(seed=0)
    N, E = x.size(0), ei.size(1)
    y = (torch.rand(N, 1, generator=torch.Generator().manual_seed(1)) < 0.1).float()
    out = {"nodes": N, "edges": E, "features": x.size(1), "cpu_cores": os.cpu_count()}
    for layers in (2, 3):
        torch.manual_seed(0)
        ref = O.OracleGAT(166, 64, 1, num_layers=layers, dropout=0.0)
        ours = GAT(166, 64, 1, num_layers=layers, dropout=0.0)
        ours.load_state_dict(ref.state_dict(), strict=True)
        ours = ours.cuda()
        xg, eg, yg = x.cuda(), ei.cuda(), y.cuda()
        crit = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor(50.0))
        critg = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor(50.0, device="cuda"))
        opt = torch.optim.Adam(ours.parameters(), lr=1e-3, weight_decay=5e-4, capturable=True)
        def fwd_gpu():
            with torch.no_grad(): ours.eval()(xg, eg)
        def step_gpu():
            ours.train(); opt.zero_grad(set_to_none=True); critg(ours(xg, eg), yg).backward(); opt.step()
        def fwd_cpu():
            with torch.no_grad(): ref.eval()(x, ei)
        def step_cpu():
            ref.train(); ref.zero_grad(set_to_none=True); crit(ref(x, ei), y).backward()
        r = {"gpu_forward_ms": gpu_time(fwd_gpu), "gpu_train_step_ms": gpu_time(step_gpu),
             "cpu_forward_ms": cpu_time(fwd_cpu), "cpu_fwd_bwd_ms": cpu_time(step_cpu)}
        try:
            r["gpu_forward_cuda_graph_ms"] = graph_time(fwd_gpu)
            r["gpu_train_step_cuda_graph_ms"] = graph_time(step_gpu)
        except Exception as e:      # capture is an optional fast path; report why it is unavailable
            r["cuda_graph_error"] = repr(e)[:300]
        r["forward_speedup"] = r["cpu_forward_ms"] / r["gpu_forward_ms"]
        r["train_speedup"] = r["cpu_fwd_bwd_ms"] / r["gpu_train_step_ms"]
        out[f"gat_{layers}layer"] = r
    torch.manual_seed(0)
    tgn = TemporalGNN(166, 64, 1, num_layers=2).cuda().eval()
    xg, eg = x.cuda(), ei.cuda()
    def tgn_fwd():
        with torch.no_grad(): tgn(xg, eg)
    out["tgn_2layer_49_snapshots_block_diagonal"] = {"gpu_forward_ms": gpu_time(tgn_fwd)}
    print(json.dumps(out))

if __name__ == "__main__":
    main()
