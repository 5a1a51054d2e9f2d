"""Snapshot builder through the C ABI vs the oracle restatement of
create_temporal_subgraph (src/data/dataset.py:198-240): integer work, so every comparison is bit-exact."""
from types import SimpleNamespace

import pytest
import torch

from gnn_fraud_detection_b200 import _abi, create_temporal_subgraph, select_steps, synth
from gnn_fraud_detection_b200.partition import snapshot_batches
from oracle import pyg_gatconv as O

pytestmark = pytest.mark.gpu


def _data(seed=0, n_steps=7, nodes_per_step=60, edges_per_step=150, cross=0):
    g = torch.Generator().manual_seed(seed)
    N = n_steps * nodes_per_step
    ts = torch.arange(N) // nodes_per_step + 1                       # Elliptic numbers its steps from 1
    perm = torch.randperm(N, generator=g)                            # interleave the steps in node-id space
    ts = ts[perm]
    ids_of = [torch.nonzero(ts == t + 1).reshape(-1) for t in range(n_steps)]
    src, dst = [], []
    for t in range(n_steps):
        k = ids_of[t]
        src.append(k[torch.randint(0, k.numel(), (edges_per_step,), generator=g)])
        dst.append(k[torch.randint(0, k.numel(), (edges_per_step,), generator=g)])
    ei = torch.stack([torch.cat(src), torch.cat(dst)])
    if cross:                                                        # edges between different steps must be dropped
        extra = torch.randint(0, N, (2, cross), generator=g)
        ei = torch.cat([ei, extra], 1)
    ei = ei[:, torch.randperm(ei.size(1), generator=g)]
    x = torch.randn(N, 5, generator=g)
    y = torch.randint(-1, 2, (N,), generator=g)
    return x, ei, ts, y


@pytest.mark.parametrize("cross", [0, 200])
def test_create_temporal_subgraph_matches_oracle(cross):
    x, ei, ts, y = _data(seed=1, cross=cross)
    d = SimpleNamespace(x=x.cuda(), edge_index=ei.cuda(), time_steps=ts.cuda(), y=y.cuda())
    for t in range(0, 9):                                            # 0 and 8 select nothing
        sub = create_temporal_subgraph(d, t)
        xo, eo, ido = O.temporal_subgraph_oracle(x, ei, ts, t)
        assert torch.equal(sub.node_indices.cpu(), ido)
        assert torch.equal(sub.edge_index.cpu(), eo)
        assert torch.equal(sub.x.cpu(), xo) and torch.equal(sub.y.cpu(), y[ido]) and torch.equal(sub.time_steps.cpu(), ts[ido])
        assert sub.edge_index.is_contiguous() and tuple(sub.edge_index.shape) == (2, eo.size(1))


def test_step_sets_equal_block_diagonal_union():
    x, ei, ts, _ = _data(seed=2)
    steps = [2, 5, 6]
    ids, sub, relabel = select_steps(ts.cuda(), ei.cuda(), steps)
    mask = torch.isin(ts, torch.tensor(steps))
    ido = torch.nonzero(mask).reshape(-1)
    rl = torch.full((ts.numel(),), -1, dtype=torch.int64)
    rl[ido] = torch.arange(ido.numel())
    em = mask[ei[0]] & mask[ei[1]]
    assert torch.equal(ids.cpu(), ido) and torch.equal(relabel.cpu(), rl) and torch.equal(sub.cpu(), rl[ei[:, em]])
    # the multi-GPU dealer goes through the same builder on CUDA tensors and through torch ops on CPU tensors
    for r in range(3):
        xc, ec, ic = snapshot_batches(x, ei, ts, r, 3)
        xg, eg, ig = snapshot_batches(x.cuda(), ei.cuda(), ts.cuda(), r, 3)
        assert torch.equal(ic, ig.cpu()) and torch.equal(ec, eg.cpu()) and torch.equal(xc, xg.cpu())


def test_edge_cases_and_errors():
    dev = torch.device("cuda")
    ts = torch.tensor([3, 3, 4], device=dev)
    e0 = torch.zeros(2, 0, dtype=torch.int64, device=dev)
    ids, sub, rl = select_steps(ts, e0, [3])                         # no edges at all
    assert ids.tolist() == [0, 1] and tuple(sub.shape) == (2, 0) and rl.tolist() == [0, 1, -1]
    ids, sub, rl = select_steps(ts, torch.tensor([[0, 2], [1, 1]], device=dev), [])   # empty selection
    assert ids.numel() == 0 and sub.size(1) == 0 and rl.tolist() == [-1, -1, -1]
    ids, sub, rl = select_steps(ts[:0], e0, [1])                     # empty graph
    assert ids.numel() == 0 and sub.size(1) == 0
    with pytest.raises(_abi.GnnfdError, match="outside"):            # endpoint out of range, as the CSR builder
        select_steps(ts, torch.tensor([[0], [7]], device=dev), [3])
    with pytest.raises(_abi.GnnfdError, match="negative time step"):
        select_steps(torch.tensor([1, -2, 1], device=dev), e0, [1])
    with pytest.raises(RuntimeError, match="CUDA"):
        select_steps(ts.cpu(), e0.cpu(), [3])


def test_elliptic_shape_partition_property():
    """Elliptic-shaped graph (49 disconnected steps): the per-step snapshots partition nodes and edges exactly."""
    x, ei, ts = synth.elliptic_synth(seed=3)
    tsg, eig = ts.cuda(), ei.cuda()
    tot_n = tot_e = 0
    seen = torch.zeros(ts.numel(), dtype=torch.bool, device="cuda")
    for t in torch.unique(ts).tolist():
        ids, sub, _ = select_steps(tsg, eig, [t])
        assert not bool(seen[ids].any())
        seen[ids] = True
        assert sub.numel() == 0 or (int(sub.min()) >= 0 and int(sub.max()) < ids.numel())
        # relabelled edges map back to original endpoints of the same step, in the original order
        orig = ids[sub]
        em = (tsg[eig[0]] == t) & (tsg[eig[1]] == t)
        assert torch.equal(orig, eig[:, em])
        tot_n += ids.numel()
        tot_e += sub.size(1)
    assert tot_n == ts.numel() and tot_e == ei.size(1) and bool(seen.all())
