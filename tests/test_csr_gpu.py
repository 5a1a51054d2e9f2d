"""GPU parity, index work: the CUDA CSR/CSC builder must be BIT-EXACT against the oracle (torch stable
sort == PyG ordering), through the C ABI."""
import numpy as np
import pytest
import torch

from gnn_fraud_detection_b200 import _abi, build_csr, synth
from oracle import pyg_gatconv as O

pytestmark = pytest.mark.gpu


def _check(ei_cpu, N, loops=True, order="dst"):
    g = build_csr(ei_cpu.cuda(), N, add_self_loops=loops, build_csc=True, order=order)
    rowptr, col, perm, ei2 = O.csr_oracle(ei_cpu, N, loops, order)
    assert g.n_edges == ei2.size(1)
    assert torch.equal(g.rowptr.cpu().long(), rowptr)
    assert torch.equal(g.col.cpu().long(), col)
    assert torch.equal(g.perm.cpu().long(), perm)
    colptr, row, eid = O.csc_oracle(rowptr, col, N)
    assert torch.equal(g.colptr.cpu().long(), colptr)
    assert torch.equal(g.csc_row.cpu().long(), row)
    assert torch.equal(g.csc_eid.cpu().long(), eid)
    return g


@pytest.mark.parametrize("N,E", [(1, 0), (5, 0), (2, 1), (33, 100), (257, 5000), (4096, 4096), (4097, 40000),
                                 (70000, 300000), (300, 70000)])
@pytest.mark.parametrize("loops", [True, False])
def test_csr_bit_exact_random(N, E, loops):
    _check(synth.random_graph(N, E, seed=N * 7 + E), N, loops)      # includes self loops and duplicates


@pytest.mark.parametrize("N,E", [(5, 0), (2, 1), (33, 100), (257, 5000), (4097, 40000), (70000, 300000), (300, 70000)])
@pytest.mark.parametrize("loops", [True, False])
def test_csr_order_dst_src_bit_exact(N, E, loops):
    """GNNFD_ORDER_DST_SRC = PyG sort_edge_index(sort_by_row=False): rows sorted by source, duplicates in input order."""
    g = _check(synth.random_graph(N, E, seed=N * 5 + E), N, loops, order="dst_src")
    rp, col = g.rowptr.cpu().long(), g.col.cpu().long()
    if g.n_edges:
        row = torch.repeat_interleave(torch.arange(N), rp[1:] - rp[:-1])
        key = row * N + col
        assert bool((key[1:] >= key[:-1]).all())


def test_layer_is_invariant_to_the_row_order():
    """The layer's result does not depend on the order of a row's edges beyond fp32 summation order."""
    from gnn_fraud_detection_b200 import GATConv
    N, E, K = 3000, 30000, 166
    ei = synth.random_graph(N, E, seed=3).cuda()
    x = torch.randn(N, K, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
    torch.manual_seed(1)
    conv = GATConv(K, 64, heads=8, concat=False).cuda().eval()
    with torch.no_grad():
        a = conv(x, build_csr(ei, N))
        b = conv(x, build_csr(ei, N, order="dst_src"))
    assert float((a - b).abs().max()) <= 2e-6


def test_csr_empty_graph():
    g = build_csr(torch.zeros(2, 0, dtype=torch.long, device="cuda"), 0)
    assert g.n_edges == 0
    g = build_csr(torch.zeros(2, 0, dtype=torch.long, device="cuda"), 4, add_self_loops=False)
    assert g.n_edges == 0 and g.rowptr.tolist() == [0, 0, 0, 0, 0]


def test_csr_all_self_loops_and_single_destination():
    N = 1000
    ar = torch.arange(N)
    _check(torch.stack([ar, ar]), N)                                  # every input edge is dropped
    _check(torch.stack([ar, torch.full((N,), 7)]), N)                 # one hub row, everything else loop-only


def test_csr_golden_fixture(golden_dir):
    import os
    gold = np.load(os.path.join(golden_dir, "golden_small.npz"))
    ei = torch.from_numpy(gold["edge_index"])
    g = build_csr(ei.cuda(), gold["x"].shape[0])
    assert np.array_equal(g.rowptr.cpu().numpy(), gold["rowptr"])
    assert np.array_equal(g.col.cpu().numpy(), gold["col"])
    assert np.array_equal(g.perm.cpu().numpy(), gold["perm"])


def test_csr_elliptic_shape_full_size():
    _, ei, _ = synth.elliptic_synth(seed=0)
    g = _check(ei, synth.ELLIPTIC_NODES)
    assert g.n_edges == synth.ELLIPTIC_EDGES + synth.ELLIPTIC_NODES   # 438,124


def test_csr_rejects_out_of_range_and_bad_dtype():
    bad = torch.tensor([[0, 9], [1, 1]], device="cuda")
    with pytest.raises(_abi.GnnfdError, match="outside"):
        build_csr(bad, 3)
    with pytest.raises(TypeError):
        build_csr(bad.int(), 10)
    with pytest.raises(RuntimeError, match="CUDA"):
        build_csr(bad.cpu(), 10)


def test_csr_large_properties():
    """size-independent properties at a size the CPU oracle would take too long for (20M keys)."""
    N, E = 2_000_000, 20_000_000
    ei = synth.powerlaw_graph(N, E, seed=3, device="cuda")
    g = build_csr(ei, N)
    Ep = g.n_edges
    dst_sorted = torch.repeat_interleave(torch.arange(N, device="cuda"), (g.rowptr[1:] - g.rowptr[:-1]).long())
    keep = ei[0] != ei[1]
    src2 = torch.cat([ei[0, keep], torch.arange(N, device="cuda")])
    dst2 = torch.cat([ei[1, keep], torch.arange(N, device="cuda")])
    perm = g.perm.long()
    assert Ep == src2.numel() and int(g.rowptr[-1]) == Ep
    assert torch.equal(torch.sort(perm).values, torch.arange(Ep, device="cuda"))          # a permutation
    assert torch.equal(dst2[perm], dst_sorted) and torch.equal(src2[perm], g.col.long())    # of the right edges
    same = dst_sorted[1:] == dst_sorted[:-1]
    assert torch.all(perm[1:][same] > perm[:-1][same])                                      # stable within a row
    # CSC: sorted by source, stable in CSR position, consistent with the CSR
    eid = g.csc_eid.long()
    srcs = g.col.long()[eid]
    assert torch.all(srcs[1:] >= srcs[:-1])
    s2 = srcs[1:] == srcs[:-1]
    assert torch.all(eid[1:][s2] > eid[:-1][s2])
    assert torch.equal(dst_sorted[eid], g.csc_row.long())
    assert torch.equal(torch.bincount(srcs, minlength=N).cumsum(0), g.colptr[1:].long())
    # hub plan covers exactly the rows above the threshold
    deg = (g.rowptr[1:] - g.rowptr[:-1]).long()
    assert g.c.hub_dst.n_hub == int((deg > _abi.HUB_THRESHOLD).sum())
    assert g.c.hub_dst.n_chunk == int(((deg[deg > _abi.HUB_THRESHOLD] + _abi.HUB_CHUNK - 1) // _abi.HUB_CHUNK).sum())
