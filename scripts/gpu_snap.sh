#!/bin/bash
timeout 240 python -m pytest tests/test_snapshot_gpu.py tests/test_models_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -12
timeout 300 python - <<'PY'
import torch, time
from gnn_fraud_detection_b200 import select_steps, synth
from gnn_fraud_detection_b200.partition import snapshot_batches
x, ei, ts = synth.elliptic_synth(seed=0)
xg, eg, tg = x.cuda(), ei.cuda(), ts.cuda()
def t(fn, n=20):
    fn(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e3
print("elliptic-size one step, CUDA builder: %.3f ms" % t(lambda: select_steps(tg, eg, [10])))
def torch_ops():
    mask = tg == 10; ids = torch.nonzero(mask).reshape(-1); rl = torch.full((tg.numel(),), -1, dtype=torch.int64, device="cuda"); rl[ids] = torch.arange(ids.numel(), device="cuda")
    em = mask[eg[0]] & mask[eg[1]]; return rl[eg[:, em]]
print("elliptic-size one step, torch index ops: %.3f ms" % t(torch_ops))
# 20M-node / 200M-edge scale
N, E = 20_000_000, 200_000_000
ts2 = torch.randint(1, 50, (N,), device="cuda"); ei2 = synth.powerlaw_graph(N, E, seed=1, device=torch.device("cuda"))
ms = t(lambda: select_steps(ts2, ei2, list(range(1, 25))), n=3)
print("20M nodes / 200M edges, 24 of 49 steps: %.1f ms  (%.0f GB/s of algorithmic traffic)" % (ms, (2*8*N + 2*16*E + 2*8*N)/ms/1e6))
PY
