import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_fraud_detection_b200 import _abi, functional as Fn
H, C, D = 8, 64, 512
def run(N, K, pts):
    x = torch.zeros(N, K); dxw = torch.zeros(N, D)
    for (n, o, k, v) in pts:
        dxw[n, o] = 1.0; x[n, k] = v
    W = torch.zeros(D, K); xw = torch.zeros(N, D); z8 = torch.zeros(N, H); dout = torch.zeros(N, C)
    args = [t.cuda() for t in (x, W, dxw, xw, z8, z8, dout)]
    dW = Fn.project_bwd(*args, H, C, C, False, _abi.GEMM_TC)[0].cpu()
    nz = dW.nonzero()
    print(f"N={N} K={K} pts={pts} -> {[(int(a), int(b), float(dW[a, b])) for a, b in nz[:12]]} (nnz={len(nz)})")
    exp = dxw.t() @ x
    print("   expected:", [(int(a), int(b), float(exp[a, b])) for a, b in exp.nonzero()[:12]])
for pts in ([(0, 0, 0, 1.0)], [(0, 1, 0, 2.0)], [(0, 0, 1, 3.0)], [(1, 0, 0, 4.0)], [(0, 5, 3, 5.0)], [(9, 40, 2, 6.0)],
            [(0, 130, 0, 7.0)], [(17, 300, 6, 8.0)], [(0, 0, 0, 1.0), (1, 0, 0, 1.0)]):
    run(32, 7, pts)
run(32, 166, [(0, 3, 40, 1.0)]); run(32, 166, [(3, 200, 165, 2.0)]); run(64, 166, [(40, 3, 100, 1.0)])
# dense check
torch.manual_seed(0)
N, K = 64, 16
x = torch.randn(N, K); dxw = torch.randn(N, D)
args = [t.cuda() for t in (x, torch.zeros(D, K), dxw, torch.zeros(N, D), torch.zeros(N, H), torch.zeros(N, H), torch.zeros(N, C))]
dW = Fn.project_bwd(*args, H, C, C, False, _abi.GEMM_TC)[0].cpu()
exp = dxw.t() @ x
print("dense: |dW|", float(dW.norm()), "|exp|", float(exp.norm()), "relerr", float((dW - exp).norm() / exp.norm()))
print(dW[:4, :8]); print(exp[:4, :8])
