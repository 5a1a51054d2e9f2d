"""Host-side mirror of the reference's model classes for the GAT hot path.

``GAT`` and ``TemporalGNN`` keep the class names, constructor arguments, ``forward`` / ``predict``
signatures, return types and ``state_dict`` keys of ``src/models/gat.py:10-122`` and
``src/models/tgn.py:14-141`` so that callers (``src/train.py:117-128``, ``src/evaluate.py:75-86``,
``compare_gnn_models.py:96-107``) and the shipped checkpoints work unchanged; the only difference is that
``GATConv`` resolves to the B200-native layer in ``gnn_fraud_detection_b200.nn``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import fused
from .nn import GATConv

__all__ = ["GAT", "TemporalGNN"]

_HEADS = 8  # hard-coded at every GATConv(...) call of the reference (gat.py:39,45,51; tgn.py:43,49,55)


class _GATStack(nn.Module):
    """The layer loop both reference models share (gat.py:78-91 == tgn.py:92-105)."""

    def __init__(self, in_channels: int, hidden_channels: int, out_channels: int, num_layers: int, dropout: float,
                 residual: bool, use_batch_norm: bool, **conv_kwargs):
        super().__init__()
        self.in_channels, self.hidden_channels, self.out_channels = in_channels, hidden_channels, out_channels
        self.num_layers, self.dropout = num_layers, dropout
        self.residual, self.use_batch_norm = residual, use_batch_norm
        # one input layer always; (num_layers-2) hidden layers; one more iff num_layers > 1
        fan_in = [in_channels] + [hidden_channels] * max(num_layers - 1, 0)
        self.gat_layers = nn.ModuleList(
            GATConv(k, hidden_channels, heads=_HEADS, concat=False, dropout=dropout, **conv_kwargs) for k in fan_in)
        self.batch_norms = (nn.ModuleList(nn.BatchNorm1d(hidden_channels) for _ in fan_in)
                            if use_batch_norm else None)
        self.fused_tail = True      # train-mode BatchNorm/ReLU/dropout/residual through libgnnfd_b200 (False: torch ops)
        self.bn_group = None        # process group whose ranks hold disjoint rows of one batch (synchronised statistics)
        self.bn_rows = None         # ... and the batch size over all its ranks, if known

    def _encode(self, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        h = x
        if not self.training and not torch.is_grad_enabled():
            # inference: BatchNorm (running stats) + ReLU + residual are fused into each layer's kernel epilogue
            for i, conv in enumerate(self.gat_layers):
                res = h if (self.residual and h.size(-1) == conv.out_channels) else None
                h = conv.forward_fused_eval(h, edge_index, self.batch_norms[i] if self.use_batch_norm else None,
                                            relu=True, residual=res)
            return h
        for i, conv in enumerate(self.gat_layers):
            z = conv(h, edge_index)
            res = h if (self.residual and h.size(-1) == z.size(-1)) else None
            if self.use_batch_norm and self.training and z.is_cuda and self.fused_tail:
                # training: BatchNorm (batch statistics) + ReLU + feature dropout + residual in two kernels (gat.py:82-91)
                h = fused.bn_relu_dropout_residual(z, self.batch_norms[i], self.dropout, res, group=self.bn_group,
                                                   global_rows=self.bn_rows)
                continue
            if self.use_batch_norm:
                z = self.batch_norms[i](z)
            z = F.dropout(F.relu(z), p=self.dropout, training=self.training)
            h = res + z if res is not None else z
        return h


class GAT(_GATStack):
    def __init__(self, in_channels: int, hidden_channels: int, out_channels: int, num_layers: int = 2,
                 dropout: float = 0.2, residual: bool = True, use_batch_norm: bool = True, **conv_kwargs):
        super().__init__(in_channels, hidden_channels, out_channels, num_layers, dropout, residual, use_batch_norm,
                         **conv_kwargs)
        self.out = nn.Linear(hidden_channels, out_channels)

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor, batch: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self.out(self._encode(x, edge_index))

    def predict(self, x, edge_index, batch=None, apply_sigmoid: bool = True) -> torch.Tensor:
        out = self.forward(x, edge_index, batch)
        return torch.sigmoid(out) if apply_sigmoid else out


class TemporalGNN(_GATStack):
    def __init__(self, in_channels: int, hidden_channels: int, out_channels: int, num_layers: int = 2,
                 dropout: float = 0.2, residual: bool = True, use_batch_norm: bool = True, **conv_kwargs):
        super().__init__(in_channels, hidden_channels, out_channels, num_layers, dropout, residual, use_batch_norm,
                         **conv_kwargs)
        self.gru = nn.GRUCell(hidden_channels, hidden_channels)
        self.out = nn.Linear(hidden_channels, out_channels)
        self.fused_head = True

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor, batch: Optional[torch.Tensor] = None,
                hidden_state: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        h = self._encode(x, edge_index)
        if h.is_cuda and self.fused_head and fused.gru_head_supported(self.gru, self.out):
            # GRUCell + Linear(64 -> 1) as one kernel; hidden_state=None is the reference's zero state (tgn.py:88-89)
            return fused.gru_head(h, self.gru, self.out, hidden_state)
        if hidden_state is None:
            hidden_state = torch.zeros(x.size(0), self.hidden_channels, device=x.device)
        hidden_state = self.gru(h, hidden_state)
        return self.out(hidden_state), hidden_state

    def predict(self, x, edge_index, batch=None, hidden_state=None, apply_sigmoid: bool = True) -> torch.Tensor:
        out, _ = self.forward(x, edge_index, batch, hidden_state)
        return torch.sigmoid(out) if apply_sigmoid else out
