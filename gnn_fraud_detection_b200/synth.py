"""Deterministic synthetic inputs shaped like the reference's data.

The Elliptic CSVs are not shipped with the reference (``elliptic_bitcoin_dataset/data.txt`` is a
placeholder), so every benchmark and parity test runs on seeded synthetic graphs with the dataset's
published shape (``processed_data/dataset_stats.json``: 203,769 nodes, 234,355 edges, 166 features,
49 time steps).  Generators run on any device; the CPU baseline receives ``.cpu()`` copies.
"""
from __future__ import annotations

import torch

ELLIPTIC_NODES, ELLIPTIC_EDGES, ELLIPTIC_FEATS, ELLIPTIC_STEPS = 203_769, 234_355, 166, 49


def _gen(device, seed: int) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def elliptic_synth(num_nodes: int = ELLIPTIC_NODES, num_edges: int = ELLIPTIC_EDGES, num_feats: int = ELLIPTIC_FEATS,
                   num_steps: int = ELLIPTIC_STEPS, seed: int = 0, device="cpu"):
    """C1/C2: ``num_steps`` disconnected time-step subgraphs, nodes numbered step-major, no self-loops,
    duplicate edges allowed, edges globally shuffled.  Returns ``x [N,F] f32, edge_index [2,E] i64,
    time_steps [N] i64 (1-based)``."""
    g = _gen(device, seed)
    w = 0.2 + 0.8 * torch.rand(num_steps, generator=g, device=device)
    n_t = torch.floor(num_nodes * w / w.sum()).long()
    n_t = torch.clamp(n_t, min=2)
    n_t[0] += num_nodes - n_t.sum()
    e_t = torch.floor(num_edges * n_t.double() / num_nodes).long()
    e_t[0] += num_edges - e_t.sum()
    start = torch.cumsum(n_t, 0) - n_t
    step_of_edge = torch.repeat_interleave(torch.arange(num_steps, device=device), e_t)
    n_e, s_e = n_t[step_of_edge], start[step_of_edge]
    dst_l = torch.floor(torch.rand(num_edges, generator=g, device=device, dtype=torch.float64) * n_e).long()
    off = 1 + torch.floor(torch.rand(num_edges, generator=g, device=device, dtype=torch.float64) * (n_e - 1)).long()
    dst_l = torch.minimum(dst_l, n_e - 1)
    off = torch.minimum(off, n_e - 1)
    src_l = (dst_l + off) % n_e                          # uniform over the step, never equal to dst
    shuffle = torch.randperm(num_edges, generator=g, device=device)
    edge_index = torch.stack([s_e + src_l, s_e + dst_l])[:, shuffle].contiguous()
    time_steps = torch.repeat_interleave(torch.arange(1, num_steps + 1, device=device), n_t)
    x = torch.randn(num_nodes, num_feats, generator=g, device=device)   # reference z-scores its features
    return x, edge_index, time_steps


def powerlaw_graph(num_nodes: int = 20_000_000, num_edges: int = 200_000_000, seed: int = 1234, device="cpu",
                   chunk: int = 1 << 26):
    """C4: ``dst = pi(floor(N*u^3))`` (in-degree density ~ rank^(-2/3), heaviest node ~ E/N^(1/3) edges),
    ``pi`` a fixed random permutation so hubs are scattered; ``src ~ U[0,N)``.  Returns edge_index [2,E]."""
    g = _gen(device, seed)
    pi = torch.randperm(num_nodes, generator=g, device=device)
    edge_index = torch.empty(2, num_edges, dtype=torch.int64, device=device)
    for b in range(0, num_edges, chunk):
        n = min(chunk, num_edges - b)
        u = torch.rand(n, generator=g, device=device, dtype=torch.float64)
        r = torch.clamp(torch.floor(num_nodes * u * u * u).long(), max=num_nodes - 1)
        edge_index[1, b:b + n] = pi[r]
        edge_index[0, b:b + n] = torch.randint(0, num_nodes, (n,), generator=g, device=device)
    return edge_index


def fraud_ring_skew(num_nodes: int = 1_000_000, background_edges: int = 5_000_000, num_hubs: int = 16,
                    hub_degree: int = 131_072, num_rings: int = 1_000, ring_len: int = 64, seed: int = 7,
                    device="cpu"):
    """C5: uniform background + ``num_hubs`` hub nodes with in-degree ``hub_degree`` (>= 1e5) + directed
    rings.  Returns edge_index [2,E] (shuffled)."""
    g = _gen(device, seed)
    parts = [torch.randint(0, num_nodes, (2, background_edges), generator=g, device=device)]
    hubs = torch.randperm(num_nodes, generator=g, device=device)[:num_hubs]
    hub_dst = hubs.repeat_interleave(hub_degree)
    hub_src = torch.randint(0, num_nodes, (num_hubs * hub_degree,), generator=g, device=device)
    parts.append(torch.stack([hub_src, hub_dst]))
    if num_rings > 0:
        members = torch.randint(0, num_nodes, (num_rings, ring_len), generator=g, device=device)
        parts.append(torch.stack([members.reshape(-1), members.roll(-1, dims=1).reshape(-1)]))
    ei = torch.cat(parts, dim=1)
    return ei[:, torch.randperm(ei.size(1), generator=g, device=device)].contiguous()


def random_graph(num_nodes: int, num_edges: int, seed: int = 0, device="cpu", self_loops: bool = True):
    """Small uniform multigraph for tests (duplicates and, optionally, self-loops occur)."""
    g = _gen(device, seed)
    ei = torch.randint(0, max(num_nodes, 1), (2, num_edges), generator=g, device=device)
    if not self_loops and num_nodes > 1:
        same = ei[0] == ei[1]
        ei[0, same] = (ei[0, same] + 1) % num_nodes
    return ei
