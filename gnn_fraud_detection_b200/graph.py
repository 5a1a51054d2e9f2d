"""Host-side graph structures: destination-sorted CSR (+CSC twin, hub plans) built by the CUDA library.

``edge_index`` is the reference's ``[2, E]`` int64 tensor (row 0 = source, row 1 = target,
``src/models/gat.py:80``).  In full-batch training the same tensor is passed to every layer of every
epoch (``src/train.py:124``), so the CSR is built once and cached on the tensor's identity
(``data_ptr``, shape, ``_version``) instead of redoing the self-loop rewrite on every forward as PyG does.
"""
from __future__ import annotations

import ctypes as C
import weakref
from collections import OrderedDict
from typing import Optional

import torch

from . import _abi


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class GraphCSR:
    """Device-resident CSR/CSC of ``edge_index'`` plus the ``gnnfd_graph_t`` view handed to the kernels."""

    def __init__(self, n_dst: int, n_src: int, n_edges: int, rowptr, col, perm, colptr=None, csc_row=None,
                 csc_eid=None, hub_threshold: int = _abi.HUB_THRESHOLD, hub_chunk: int = _abi.HUB_CHUNK,
                 plan_hubs: bool = True):
        self.n_dst, self.n_src, self.n_edges = int(n_dst), int(n_src), int(n_edges)
        self.rowptr, self.col, self.perm = rowptr, col, perm
        self.colptr, self.csc_row, self.csc_eid = colptr, csc_row, csc_eid
        self.device = rowptr.device
        self._hub_tensors = []
        self.c = _abi.Graph()
        self.c.n_dst, self.c.n_src, self.c.n_edges = self.n_dst, self.n_src, self.n_edges
        self.c.rowptr, self.c.col, self.c.perm = _abi.ptr(rowptr), _abi.ptr(col), _abi.ptr(perm)
        self.c.colptr, self.c.csc_row, self.c.csc_eid = _abi.ptr(colptr), _abi.ptr(csc_row), _abi.ptr(csc_eid)
        self.csr2csc = None
        if csc_eid is not None and self.n_edges > 0:
            self.csr2csc = torch.empty(self.n_edges, dtype=torch.int32, device=self.device)
            _abi.check(_abi.lib().gnnfd_invert_perm(csc_eid.data_ptr(), self.n_edges, self.csr2csc.data_ptr(), _stream()))
        self.c.csr2csc = _abi.ptr(self.csr2csc)
        if plan_hubs:
            self.c.hub_dst = self._plan(rowptr, self.n_dst, hub_threshold, hub_chunk)
            if colptr is not None:
                self.c.hub_src = self._plan(colptr, self.n_src, hub_threshold, hub_chunk)
        # per-row charge in edge units: a dst row costs ~one 256 B epilogue, a src row writes a 2 KB dxw row
        self.c.items_dst = self._items(rowptr, self.n_dst, row_weight=1)
        self._item_start_dst = self._hub_tensors[-1] if self.c.items_dst.n_items > 0 else None
        self._item_blocks = {}
        if colptr is not None:
            self.c.items_src = self._items(colptr, self.n_src, row_weight=8)

    @property
    def has_csc(self) -> bool:
        return self.colptr is not None

    def _items(self, ptr_t, n_rows, row_weight=1) -> _abi.ItemPlan:
        """Cost-balanced work items: ~target (edges + rows*row_weight) per warp, sized so that even a small graph
        yields several items per resident warp (148 SMs x 32 warps)."""
        plan = _abi.ItemPlan()
        cost = self.n_edges + n_rows * row_weight
        target = int(min(256, max(32, (cost // (148 * 32 * 4)) // 32 * 32)))
        plan.target = target
        if n_rows == 0:
            return plan
        n_items = cost // target + 1
        item_start = torch.empty(n_items + 1, dtype=torch.int32, device=self.device)
        _abi.check(_abi.lib().gnnfd_item_plan(ptr_t.data_ptr(), n_rows, self.n_edges, target, row_weight,
                                              item_start.data_ptr(), _stream()))
        self._hub_tensors.append(item_start)
        plan.n_items, plan.item_start = n_items, item_start.data_ptr()
        return plan

    def _plan(self, ptr_t, n_rows, threshold, chunk) -> _abi.HubPlan:
        plan = _abi.HubPlan()
        plan.threshold, plan.chunk = threshold, chunk
        if n_rows == 0 or self.n_edges <= threshold:
            return plan
        L = _abi.lib()
        cap_hub = self.n_edges // (threshold + 1) + 1
        cap_chunk = self.n_edges // chunk + cap_hub + 1
        hub_row = torch.empty(cap_hub, dtype=torch.int32, device=self.device)
        hub_chunk_ptr = torch.empty(cap_hub + 1, dtype=torch.int32, device=self.device)
        chunk_hub = torch.empty(cap_chunk, dtype=torch.int32, device=self.device)
        nbytes = C.c_size_t()
        _abi.check(L.gnnfd_hub_plan_workspace_bytes(n_rows, C.byref(nbytes)))
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
        counts = (C.c_int64 * 2)()
        _abi.check(L.gnnfd_hub_plan(ptr_t.data_ptr(), n_rows, threshold, chunk, hub_row.data_ptr(),
                                    hub_chunk_ptr.data_ptr(), chunk_hub.data_ptr(), cap_hub, cap_chunk, counts,
                                    ws.data_ptr(), nbytes.value, _stream()))
        n_hub, n_chunk = int(counts[0]), int(counts[1])
        if n_hub == 0:
            return plan
        hub_row, hub_chunk_ptr, chunk_hub = hub_row[:n_hub].clone(), hub_chunk_ptr[:n_hub + 1].clone(), \
            chunk_hub[:n_chunk].clone()
        self._hub_tensors += [hub_row, hub_chunk_ptr, chunk_hub]
        plan.n_hub, plan.n_chunk = n_hub, n_chunk
        plan.hub_row, plan.hub_chunk_ptr, plan.chunk_hub = hub_row.data_ptr(), hub_chunk_ptr.data_ptr(), \
            chunk_hub.data_ptr()
        return plan

    def item_blocks(self, n_blocks: int):
        """Split the dst-major work items into ``n_blocks`` contiguous runs: [(item_lo, item_hi, row_lo, row_hi)].
        Item boundaries are row boundaries, so a block is a contiguous range of destination rows (cached; one small
        device->host read per graph and block count)."""
        n_items = int(self.c.items_dst.n_items)
        if n_items == 0:
            return [(0, 0, 0, self.n_dst)]
        n_blocks = max(1, min(int(n_blocks), n_items))
        hit = self._item_blocks.get(n_blocks)
        if hit is None:
            cuts = [b * n_items // n_blocks for b in range(n_blocks + 1)]
            rows = self._item_start_dst[torch.tensor(cuts, device=self.device)].tolist()
            rows[0], rows[-1] = 0, self.n_dst
            hit = [(cuts[b], cuts[b + 1], int(rows[b]), int(rows[b + 1])) for b in range(n_blocks)]
            self._item_blocks[n_blocks] = hit
        return hit

    def ref(self):
        return C.byref(self.c)


def build_csr(edge_index: torch.Tensor, num_nodes: int, add_self_loops: bool = True, build_csc: bool = True,
              hub_threshold: int = _abi.HUB_THRESHOLD, hub_chunk: int = _abi.HUB_CHUNK, order: str = "dst") -> GraphCSR:
    """edge_index [2,E] int64 (cuda) -> GraphCSR, via ``gnnfd_csr_build`` (stable radix sort on device).
    ``order="dst"``: edges of a row in their order of appearance (what PyG's scatter sees, the default);
    ``order="dst_src"``: additionally sorted by source within a row (PyG ``sort_edge_index(sort_by_row=False)``)."""
    if order not in ("dst", "dst_src"):
        raise ValueError(f"order must be 'dst' or 'dst_src', got {order!r}")
    if edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise ValueError(f"edge_index must be [2, E], got {tuple(edge_index.shape)}")
    if edge_index.dtype != torch.int64:
        raise TypeError(f"edge_index must be int64 (torch.long), got {edge_index.dtype}")
    if not edge_index.is_cuda:
        raise RuntimeError("edge_index must live on a CUDA device: this path has no CPU implementation")
    L = _abi.lib()
    ei = edge_index.contiguous()
    E, N, dev = ei.size(1), int(num_nodes), ei.device
    flags = ((_abi.ADD_SELF_LOOPS if add_self_loops else 0) | (_abi.BUILD_CSC if build_csc else 0)
             | (_abi.ORDER_DST_SRC if order == "dst_src" else 0))
    cap = E + (N if add_self_loops else 0)
    with torch.cuda.device(dev):
        nbytes = C.c_size_t()
        _abi.check(L.gnnfd_csr_workspace_bytes(N, E, flags, C.byref(nbytes)))
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        rowptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
        col = torch.empty(cap, dtype=torch.int32, device=dev)
        perm = torch.empty(cap, dtype=torch.int32, device=dev)
        colptr = csc_row = csc_eid = None
        if build_csc:
            colptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
            csc_row = torch.empty(cap, dtype=torch.int32, device=dev)
            csc_eid = torch.empty(cap, dtype=torch.int32, device=dev)
        e_out = C.c_int64()
        _abi.check(L.gnnfd_csr_build(ei.data_ptr(), E, N, flags, rowptr.data_ptr(), col.data_ptr(), perm.data_ptr(),
                                     _abi.ptr(colptr), _abi.ptr(csc_row), _abi.ptr(csc_eid), C.byref(e_out),
                                     ws.data_ptr(), nbytes.value, _stream()))
        Ep = int(e_out.value)
        del ws
        if Ep != cap:  # existing self-loops were dropped: trim the over-allocated tails
            col, perm = col[:Ep].clone(), perm[:Ep].clone()
            if build_csc:
                csc_row, csc_eid = csc_row[:Ep].clone(), csc_eid[:Ep].clone()
        return GraphCSR(N, N, Ep, rowptr, col, perm, colptr, csc_row, csc_eid, hub_threshold, hub_chunk)


class CSRCache:
    """Small LRU keyed on the identity of the edge_index tensor (full-batch training reuses it).

    The key is ``(data_ptr, shape, strides, _version, device, num_nodes, add_self_loops)``.  An entry does NOT keep the
    tensor alive: it holds a weak reference and a finalizer drops the entry when the tensor dies, so a one-shot
    ``edge_index`` (the reference's train loop moves a fresh batch to the device every step, ``src/train.py:105``) does not
    pin its CSR/CSC in HBM, and a recycled ``data_ptr`` cannot hit a stale entry.  Inference tensors carry no version
    counter and are built uncached.  For streaming use, build the graph once with ``build_csr`` and pass the ``GraphCSR``.
    """

    def __init__(self, capacity: int = 8):
        self.capacity = capacity
        self._d: "OrderedDict[tuple, tuple]" = OrderedDict()

    def get(self, edge_index: torch.Tensor, num_nodes: int, add_self_loops: bool, build_csc: bool) -> GraphCSR:
        if edge_index.is_inference():
            return build_csr(edge_index, num_nodes, add_self_loops, build_csc)
        key = (edge_index.data_ptr(), tuple(edge_index.shape), tuple(edge_index.stride()), edge_index._version,
               edge_index.device.index, int(num_nodes), bool(add_self_loops))
        hit = self._d.get(key)
        if hit is not None and hit[1]() is edge_index and (hit[0].has_csc or not build_csc):
            self._d.move_to_end(key)
            return hit[0]
        g = build_csr(edge_index, num_nodes, add_self_loops, build_csc)
        ref = weakref.ref(edge_index)
        weakref.finalize(edge_index, self._d.pop, key, None)
        self._d[key] = (g, ref)
        self._d.move_to_end(key)
        while len(self._d) > self.capacity:
            self._d.popitem(last=False)
        return g

    def clear(self):
        self._d.clear()


GLOBAL_CSR_CACHE = CSRCache()
