"""CPU tests of the host-side mirror of the reference interface (no kernels run)."""
import os

import numpy as np
import pytest
import torch

import gnn_fraud_detection_b200 as pkg
from gnn_fraud_detection_b200 import synth
from oracle import pyg_gatconv as O


def _ckpt(path):
    npz = np.load(path)
    sd = {k: torch.from_numpy(npz[k]) for k in npz.files}
    for k in list(sd):
        if k.endswith("lin_src.weight"):
            sd[k.replace("lin_src", "lin_dst")] = sd[k]
    return sd


def test_gatconv_state_dict_keys_match_pyg_2x():
    conv = pkg.GATConv(166, 64, heads=8, concat=False, dropout=0.2)
    keys = set(conv.state_dict().keys())
    assert keys == {"att_src", "att_dst", "bias", "lin_src.weight", "lin_dst.weight"}
    assert conv.lin_dst is conv.lin_src
    assert tuple(conv.att_src.shape) == (1, 8, 64) and tuple(conv.bias.shape) == (64,)
    assert tuple(conv.lin_src.weight.shape) == (512, 166)
    assert len(list(conv.parameters())) == 4                      # the alias is not double counted
    assert tuple(pkg.GATConv(10, 64, heads=8, concat=True).bias.shape) == (512,)
    assert torch.all(conv.bias == 0)


def test_reference_checkpoints_load_strict(golden_dir):
    gat = pkg.GAT(165, 64, 1, num_layers=3)
    gat.load_state_dict(_ckpt(os.path.join(golden_dir, "gat_ckpt.npz")), strict=True)
    tgn = pkg.TemporalGNN(165, 64, 1, num_layers=3)
    tgn.load_state_dict(_ckpt(os.path.join(golden_dir, "tgn_ckpt.npz")), strict=True)
    # same key set as the oracle's restatement of the reference classes
    assert set(gat.state_dict()) == set(O.OracleGAT(165, 64, 1, num_layers=3).state_dict())
    assert set(tgn.state_dict()) == set(O.OracleTemporalGNN(165, 64, 1, num_layers=3).state_dict())


@pytest.mark.parametrize("num_layers,expect", [(1, 1), (2, 2), (3, 3), (5, 5)])
def test_layer_construction_rule(num_layers, expect):
    m = pkg.GAT(166, 64, 1, num_layers=num_layers)
    assert len(m.gat_layers) == expect and len(m.batch_norms) == expect
    assert m.gat_layers[0].in_channels == 166
    assert all(l.in_channels == 64 for l in m.gat_layers[1:])
    assert all(l.heads == 8 and not l.concat for l in m.gat_layers)


def test_unsupported_pyg_features_raise():
    with pytest.raises(NotImplementedError):
        pkg.GATConv((4, 4), 8)
    with pytest.raises(NotImplementedError):
        pkg.GATConv(4, 8, edge_dim=3)
    conv = pkg.GATConv(4, 64, heads=8, concat=False)
    with pytest.raises(RuntimeError, match="CUDA"):
        conv(torch.randn(5, 4), torch.zeros(2, 0, dtype=torch.long))      # CPU tensor: no fallback


def test_missing_extension_fails_loudly(monkeypatch):
    from gnn_fraud_detection_b200 import _abi
    monkeypatch.setattr(_abi, "_lib", None)
    monkeypatch.setattr(_abi, "LIB_PATH", "/nonexistent/libgnnfd_b200.so")
    with pytest.raises(ImportError, match="no CPU/PyTorch fallback"):
        _abi.lib()


def test_product_path_does_not_import_the_oracle():
    root = os.path.dirname(os.path.abspath(pkg.__file__))
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_elliptic_synth_shape_and_structure():
    x, ei, ts = synth.elliptic_synth(num_nodes=5000, num_edges=6000, num_feats=166, seed=0)
    assert x.shape == (5000, 166) and ei.shape == (2, 6000) and ts.shape == (5000,)
    assert ei.dtype == torch.int64 and int(ts.min()) == 1 and int(ts.max()) == 49
    assert torch.all(ts[ei[0]] == ts[ei[1]])                      # no edge crosses a time step
    assert torch.all(ei[0] != ei[1])
    x2, ei2, _ = synth.elliptic_synth(num_nodes=5000, num_edges=6000, num_feats=166, seed=0)
    assert torch.equal(ei, ei2) and torch.equal(x, x2)


def test_skew_and_powerlaw_generators():
    ei = synth.fraud_ring_skew(num_nodes=20000, background_edges=50000, num_hubs=3, hub_degree=4096, num_rings=10,
                               ring_len=8, seed=7)
    deg = torch.bincount(ei[1], minlength=20000)
    assert int((deg >= 4096).sum()) == 3
    pl = synth.powerlaw_graph(num_nodes=50000, num_edges=500000, seed=1)
    assert pl.shape == (2, 500000) and int(pl.max()) < 50000 and int(pl.min()) >= 0
    assert int(torch.bincount(pl[1], minlength=50000).max()) > 5000      # ~ E / N^(1/3)


def test_snapshot_builder_has_no_cpu_fallback_and_validates_arguments():
    """create_temporal_subgraph / select_steps (src/data/dataset.py:198-240 mirror) run in the CUDA library only."""
    from types import SimpleNamespace

    from gnn_fraud_detection_b200 import create_temporal_subgraph, select_steps
    ts = torch.tensor([1, 1, 2])
    ei = torch.tensor([[0, 2], [1, 2]])
    with pytest.raises(RuntimeError, match="CUDA"):
        select_steps(ts, ei, [1])
    with pytest.raises(RuntimeError, match="CUDA"):
        create_temporal_subgraph(SimpleNamespace(x=torch.zeros(3, 2), edge_index=ei, time_steps=ts, y=None), 1)
    # the multi-GPU dealer keeps a torch-ops path for CPU tensors (host-side planning, gloo tests): same semantics
    from gnn_fraud_detection_b200.partition import snapshot_batches
    from oracle import pyg_gatconv as O
    x = torch.arange(6.0).view(3, 2)
    xl, el, ids = snapshot_batches(x, ei, ts, 0, 1)
    assert torch.equal(ids, torch.tensor([0, 1, 2])) and torch.equal(el, ei) and torch.equal(xl, x)
    xo, eo, ido = O.temporal_subgraph_oracle(x, ei, ts, 1)
    assert ido.tolist() == [0, 1] and eo.tolist() == [[0], [1]]


# ---- CSR cache: key, lifetime, inference tensors -----------------------------------------------------
class _FakeGraph:
    has_csc = True


def test_csr_cache_key_lifetime_and_inference_tensors(monkeypatch):
    import gc
    from gnn_fraud_detection_b200 import graph
    built = []

    def fake_build(edge_index, num_nodes, add_self_loops=True, build_csc=True, **kw):
        built.append((edge_index.data_ptr(), tuple(edge_index.shape), tuple(edge_index.stride())))
        return _FakeGraph()

    monkeypatch.setattr(graph, "build_csr", fake_build)
    cache = graph.CSRCache(capacity=4)
    base = torch.arange(24, dtype=torch.int64).view(2, 12)
    a = base[:, :3]                                     # same data_ptr / shape / version, different strides
    b = base.view(-1)[:6].view(2, 3)
    ga, gb = cache.get(a, 5, True, True), cache.get(b, 5, True, True)
    assert ga is not gb and len(built) == 2
    assert cache.get(a, 5, True, True) is ga and len(built) == 2          # hit
    base[0, 0] = 7                                                        # in-place edit bumps _version: rebuilt
    assert cache.get(a, 5, True, True) is not ga and len(built) == 3
    # an entry does not keep its tensor alive, and dies with it
    t = torch.zeros(2, 4, dtype=torch.int64)
    cache.get(t, 9, True, True)
    n = len(cache._d)
    del t
    gc.collect()
    assert len(cache._d) == n - 1
    # inference tensors have no version counter: built uncached instead of raising
    with torch.inference_mode():
        u = torch.zeros(2, 4, dtype=torch.int64)
        n_built, n_entries = len(built), len(cache._d)
        cache.get(u, 9, True, True)
        cache.get(u, 9, True, True)
    assert len(built) == n_built + 2 and len(cache._d) == n_entries
