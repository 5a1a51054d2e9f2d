"""GPU checks of the module-level API contract: argument validation on every path that hands raw pointers to the C ABI
(ADVICE round 1), the dropout hooks, and the CSR cache with real graphs."""
import pytest
import torch

from gnn_fraud_detection_b200 import GAT, GATConv, _abi
from gnn_fraud_detection_b200.graph import GLOBAL_CSR_CACHE

pytestmark = pytest.mark.gpu


def _ei(n=50, e=200, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, n, (2, e), generator=g).cuda()


@pytest.mark.parametrize("algo", [_abi.GEMM_AUTO, _abi.GEMM_INPUT])
def test_fused_eval_path_validates_like_the_training_path(algo):
    conv = GATConv(16, 64, heads=8, concat=False, gemm_algo=algo).cuda().eval()
    ei = _ei()
    x = torch.randn(50, 16, device="cuda")
    conv.forward_fused_eval(x, ei)                                           # fine
    with pytest.raises(TypeError):
        conv.forward_fused_eval(x.double(), ei)                              # fp64 would be read as fp32
    with pytest.raises(TypeError):
        conv.forward_fused_eval(x.half(), ei)                                # fp16 would be read out of bounds
    with pytest.raises(ValueError):
        conv.forward_fused_eval(torch.randn(50, 12, device="cuda"), ei)      # wrong feature count
    with pytest.raises(ValueError):
        conv.forward_fused_eval(torch.randn(40, 16, device="cuda"), conv._graph(ei, 50))   # fewer rows than the graph
    with pytest.raises(ValueError):
        conv.forward_fused_eval(x, ei, residual=torch.randn(50, 32, device="cuda"))
    with pytest.raises(RuntimeError):
        conv.forward_fused_eval(x.cpu(), ei)
    # the model-level eval path goes through the same checks
    gat = GAT(16, 64, 1, num_layers=2).cuda().eval()
    with torch.no_grad(), pytest.raises(TypeError):
        gat(x.double(), ei)


def test_dropout_hooks():
    x, ei = torch.randn(50, 16, device="cuda"), _ei()
    conv0 = GATConv(16, 64, heads=8, concat=False, dropout=0.0).cuda()
    with pytest.raises(ValueError):                      # a mask that would silently be ignored
        conv0(x, ei, dropout_mask=torch.ones(250, 8, dtype=torch.bool))
    conv1 = GATConv(16, 64, heads=8, concat=False, dropout=1.0).cuda().train()
    with torch.no_grad():
        conv1.bias.fill_(0.25)
    out = conv1(x, ei)                                   # PyG: every coefficient dropped => bias only
    assert torch.equal(out, torch.full_like(out, 0.25))
    conv1.eval()
    assert float((conv1(x, ei) - 0.25).abs().max()) > 1e-3


def test_csr_cache_views_and_inference_mode():
    GLOBAL_CSR_CACHE.clear()
    conv = GATConv(16, 64, heads=8, concat=False).cuda().eval()
    x = torch.randn(30, 16, device="cuda")
    base = torch.randint(0, 30, (2, 12), device="cuda")
    a, b = base[:, :3], base.view(-1)[:6].view(2, 3)     # same data_ptr, shape and version; different strides
    with torch.no_grad():
        oa, ob = conv(x, a), conv(x, b)
        assert torch.allclose(oa, conv(x, a.clone())) and torch.allclose(ob, conv(x, b.clone()))
        assert not torch.equal(oa, ob) or torch.equal(a, b)
    with torch.inference_mode():
        ei = torch.randint(0, 30, (2, 40), device="cuda")
        o1 = conv(x, ei)
        assert torch.equal(o1, conv(x, ei))
    GLOBAL_CSR_CACHE.clear()
