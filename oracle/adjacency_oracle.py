"""CPU oracle for the adjacency / batch index bookkeeping.  TEST INFRASTRUCTURE ONLY.

Restates, with numpy array ops, the index arithmetic that defines the layout the hot path
consumes:
  * process_adjacency_lists            /root/reference/tf2_gnn/data/utils.py:9-58
      - backward edges, tied or fresh type  utils.py:102-113
      - self loops at a chosen type slot    utils.py:91-99, 36-52
      - in-degree table [L, V]              utils.py:116-124
  * disjoint-graph batch assembly      /root/reference/tf2_gnn/data/graph_dataset.py:161-246

PINNED: tests/golden/process_adjacency_lists_golden.json is produced by running the
reference's own tf2_gnn/data/utils.py (pure numpy, importable without TensorFlow) through
oracle/gen_golden.py, and also contains the 8 expected outputs transcribed from
tf2_gnn/test/data/test_utils.py:50-115.  Bit-exact (np.array_equal) parity is required.
assemble_batch is PINNED the same way: tests/golden/batch_assembly_golden.json holds minibatches produced by
executing the reference's own tf2_gnn/data/graph_dataset.py:161-246 on seeded random graphs (oracle/gen_golden.py
imports that module with tensorflow / dpu_utils stubbed: they are only used by its tf.data wrapper).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Set, Tuple, Union

import numpy as np


def get_tied_edge_types(tie_fwd_bkwd_edges: Union[bool, List[int]], num_fwd_edge_types: int) -> Set[int]:
    """utils.py:61-77."""
    if isinstance(tie_fwd_bkwd_edges, list):
        return set(tie_fwd_bkwd_edges)
    return set(range(num_fwd_edge_types)) if tie_fwd_bkwd_edges else set()


def compute_number_of_edge_types(tied: Set[int], num_fwd_edge_types: int, add_self_loop_edges: bool) -> int:
    """utils.py:80-84."""
    return 2 * num_fwd_edge_types - len(tied) + int(add_self_loop_edges)


def _as_pairs(adj) -> np.ndarray:
    a = np.asarray(adj, dtype=np.int32)
    return a.reshape(-1, 2) if a.size else np.zeros((0, 2), dtype=np.int32)


def process_adjacency_lists(adjacency_lists: Sequence, num_nodes: int, add_self_loop_edges: bool,
                            tied_fwd_bkwd_edge_types: Set[int], self_loop_edge_type: int = 0
                            ) -> Tuple[List[np.ndarray], np.ndarray]:
    fwd = [_as_pairs(a) for a in adjacency_lists]
    out: List[np.ndarray] = [a.copy() for a in fwd]
    fresh: List[np.ndarray] = []
    for t, a in enumerate(fwd):
        flipped = a[:, ::-1]
        if t in tied_fwd_bkwd_edge_types:
            out[t] = np.concatenate([out[t], flipped], axis=0)
        else:
            fresh.append(flipped.copy())
    out.extend(fresh)
    if add_self_loop_edges:
        n_types = len(out)
        if not (-(n_types + 1) <= self_loop_edge_type <= n_types):
            raise AssertionError(
                f"Self loop edge type {self_loop_edge_type} should be in range "
                f"[{-(n_types + 1)}, {n_types}].")
        slot = self_loop_edge_type + n_types + 1 if self_loop_edge_type < 0 else self_loop_edge_type
        ids = np.arange(num_nodes, dtype=np.int32)
        out.insert(slot, np.stack([ids, ids], axis=1))
    counts = np.zeros((len(out), num_nodes), dtype=np.float64)  # reference dtype: np.zeros default
    for t, a in enumerate(out):
        if len(a):
            counts[t] = np.bincount(a[:, 1], minlength=num_nodes)
    return [np.ascontiguousarray(a, dtype=np.int32) for a in out], counts


def assemble_batch(graphs: Sequence[Dict], num_edge_types: int) -> Dict[str, np.ndarray]:
    """graph_dataset.py:204-246: concatenate graphs into one disjoint graph.  Each graph is
    {"node_features": [n, F], "adjacency_lists": [L x [e,2]]}; node ids are offset by the
    running node count (:218-222) and node_to_graph_map is a constant block per graph
    (:211-217); empty types become int32[0,2] (:244)."""
    feats, n2g, adj = [], [], [[] for _ in range(num_edge_types)]
    offset = 0
    for g_idx, g in enumerate(graphs):
        nf = np.asarray(g["node_features"])
        n = len(nf)
        feats.append(nf)
        n2g.append(np.full((n,), g_idx, dtype=np.int32))
        for t in range(num_edge_types):
            adj[t].append(_as_pairs(g["adjacency_lists"][t]) + np.int32(offset))
        offset += n
    batch = {
        "node_features": np.concatenate(feats, axis=0) if feats else np.zeros((0, 0), np.float32),
        "node_to_graph_map": np.concatenate(n2g) if n2g else np.zeros((0,), np.int32),
        "num_graphs_in_batch": len(graphs),
    }
    for t in range(num_edge_types):
        batch[f"adjacency_list_{t}"] = (np.concatenate(adj[t], axis=0).astype(np.int32)
                                        if adj[t] else np.zeros((0, 2), np.int32))
    return batch
