"""Snapshot builder: the step immediately before the hot path for the temporal configuration.

``create_temporal_subgraph(data, time_step)`` keeps the name, arguments and result attributes of the reference's
function (``src/data/dataset.py:198-240``): the nodes of one time step (ascending original id), the edges with both
endpoints inside (original order), endpoints relabelled ``0..n_t-1``, and ``x`` / ``y`` / ``time_steps`` restricted to
those nodes.  The reference walks the edge list in a Python loop over 0-d tensors (minutes on Elliptic, and its
``src in idx_mapping`` test compares tensors with int keys, so it keeps no edges -- dead code there); here the
selection, relabelling and order-preserving edge compaction are two prefix-sum compactions in ``libgnnfd_b200.so``
(``gnnfd_subgraph_build``).  ``select_steps`` is the same builder for a SET of time steps (the block-diagonal batch
that ``partition.snapshot_batches`` deals to a GPU).  CUDA tensors only; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Iterable, Tuple

import torch

from . import _abi

__all__ = ["create_temporal_subgraph", "select_steps"]


def select_steps(time_steps: torch.Tensor, edge_index: torch.Tensor, steps: Iterable[int]
                 ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """``(node_ids [n_sel] int64, edge_index_sub [2,m] int64, relabel [N] int64)`` for the nodes whose time step is
    in ``steps``; orders and numbering as in ``create_temporal_subgraph``."""
    if not (time_steps.is_cuda and edge_index.is_cuda):
        raise RuntimeError("gnn_fraud_detection_b200.snapshot runs on CUDA tensors only (no CPU fallback)")
    if time_steps.dtype != torch.int64 or edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise ValueError("time_steps must be int64 [N] and edge_index int64 [2,E]")
    lib = _abi.lib()
    dev = time_steps.device
    N, E = time_steps.numel(), edge_index.size(1)
    steps = [int(t) for t in steps]
    if any(t < 0 for t in steps):
        raise ValueError("time steps must be non-negative")
    # the table only has to reach the largest SELECTED step: larger values in the data are simply not selected
    n_table = max([t + 1 for t in steps], default=0)
    host_table = torch.zeros(max(n_table, 1), dtype=torch.uint8)
    if steps:
        host_table[torch.tensor(steps, dtype=torch.long)] = 1
    table = host_table.to(dev)
    ts = time_steps.contiguous()
    ei = edge_index.contiguous()
    node_ids = torch.empty(N, dtype=torch.int64, device=dev)
    relabel = torch.empty(N, dtype=torch.int64, device=dev)
    sub = torch.empty(2, E, dtype=torch.int64, device=dev)
    need = C.c_size_t(0)
    _abi.check(lib.gnnfd_subgraph_workspace_bytes(N, E, C.byref(need)))
    ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
    counts = (C.c_int64 * 2)()
    with torch.cuda.device(dev):
        _abi.check(lib.gnnfd_subgraph_build(ts.data_ptr(), N, table.data_ptr(), n_table, ei.data_ptr(), E,
                                            node_ids.data_ptr(), relabel.data_ptr(), sub.data_ptr(), E, counts,
                                            ws.data_ptr(), need.value, torch.cuda.current_stream(dev).cuda_stream))
    n_sel, m = int(counts[0]), int(counts[1])
    return node_ids[:n_sel], sub[:, :m].contiguous(), relabel


def create_temporal_subgraph(data, time_step: int):
    """Drop-in for ``src/data/dataset.py:198`` on any object with ``x``, ``edge_index``, ``time_steps`` (and
    optionally ``y``) CUDA tensors; returns an object with the same attribute names (plus ``node_indices``)."""
    node_ids, ei, _ = select_steps(data.time_steps, data.edge_index, [int(time_step)])
    out = SimpleNamespace(x=data.x[node_ids], edge_index=ei, time_steps=data.time_steps[node_ids], node_indices=node_ids)
    y = getattr(data, "y", None)
    out.y = y[node_ids] if y is not None else None
    return out
