"""Timed CPU baseline: the reference's op sequence restated on torch-CPU.  TEST/BENCH INFRASTRUCTURE.

TensorFlow and dpu_utils are not installed in this image (nor on the GPU box), so "the reference's
own TF2 CPU path" cannot be executed.  This module mirrors its materialisation pattern op for op
with multi-threaded torch-CPU kernels (MKL matmul, index_select, index_add_), which is what
bench.py's `cpu_baseline` and `--impl reference` legs time ("kind": "port"):

  per edge type:  index_select(h, src) -> index_select(h, tgt) -> mm(W_l) -> * 1/(c+1e-7)
  (message_passing.py:181-218, gnn_edge_mlp.py:84-107), then cat over types -> index_add_ into
  [V,H] -> activation (message_passing.py:166-177); the in-degree table is recomputed per layer as
  the reference does (message_passing.py:190,252-263).

It is validated against oracle/message_passing_oracle.py in tests/test_oracle_golden.py.
Only bench.py and tests/ may import this module.
"""
from __future__ import annotations

from typing import Sequence

import torch


def rgcn_layer_reference_order(h: torch.Tensor, adjacency_lists: Sequence[torch.Tensor],
                               weights: Sequence[torch.Tensor], normalize: bool = True,
                               activation: str = "relu", materialise_target_states: bool = True) -> torch.Tensor:
    V = h.shape[0]
    # calculate_type_to_num_incoming_edges: scatter_nd of ones per type (message_passing.py:252-263)
    counts = []
    for adj in adjacency_lists:
        c = torch.zeros(V, dtype=torch.float32)
        c.index_add_(0, adj[:, 1].long(), torch.ones(adj.shape[0], dtype=torch.float32))
        counts.append(c)
    messages, targets = [], []
    for l, adj in enumerate(adjacency_lists):
        src = adj[:, 0].long()
        tgt = adj[:, 1].long()
        edge_source_states = h.index_select(0, src)                     # message_passing.py:197-199
        if materialise_target_states:
            _edge_target_states = h.index_select(0, tgt)                # :200-202 (eager TF gathers it)
        n_in = counts[l].index_select(0, tgt)                           # :204-206
        m = edge_source_states @ weights[l]                             # gnn_edge_mlp.py:100
        if normalize:
            m = (1.0 / (n_in + 1e-7)).unsqueeze(-1) * m                 # gnn_edge_mlp.py:102-106
        messages.append(m)
        targets.append(tgt)
    all_messages = torch.cat(messages, dim=0)                           # message_passing.py:166-167
    all_targets = torch.cat(targets, dim=0)
    out = torch.zeros((V, weights[0].shape[1]), dtype=torch.float32)
    out.index_add_(0, all_targets, all_messages)                        # unsorted_segment_sum :172-174
    if activation == "relu":
        out = torch.relu(out)                                           # :176-177
    elif activation == "tanh":
        out = torch.tanh(out)
    elif activation is not None:
        raise ValueError(activation)
    return out
