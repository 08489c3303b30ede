"""A dataset fold packed on the device, and minibatches assembled from it by CUDA kernels.

Mirrors ``GraphDataset.graph_batch_iterator_from_graph_iterator`` / ``_add_graph_to_batch`` / ``_finalise_batch``
(tf2_gnn/data/graph_dataset.py:161-246): a minibatch is the disjoint union of some graphs, node ids offset by the
running node count, ``node_to_graph_map`` a constant block per graph, empty edge types ``int32[0, 2]``.  The
reference rebuilds these arrays with Python loops for every batch; here the graphs are uploaded ONCE (all graphs of
an edge type back to back, graph-local ids) and ``batch(graph_ids)`` launches ``tfgnn_b200_assemble_batch``
(batch_builder.cu).  Only the greedy "which graphs fit" rule (:181-188) stays on the host: it needs node counts only.
"""
from __future__ import annotations

import ctypes
from ctypes import c_int64, c_void_p
from typing import Any, Dict, Iterator, List, Optional, Sequence

import numpy as np
import torch

from .. import _ffi
from ..runtime import require_cuda, stream_ptr


class DeviceGraphStore:
    """graphs: sequence of samples with ``node_features`` ([n, F] array-like) and ``adjacency_lists`` (list of
    num_edge_types arrays reshapeable to [e, 2], graph-local node ids), as ``GraphSample`` in graph_dataset.py:17-41."""

    def __init__(self, graphs: Sequence[Any], num_edge_types: int):
        self.device = require_cuda()
        self.num_edge_types = int(num_edge_types)
        feats, node_counts = [], []
        edges: List[List[np.ndarray]] = [[] for _ in range(self.num_edge_types)]
        edge_counts = np.zeros((self.num_edge_types, len(graphs)), dtype=np.int64)
        for g, sample in enumerate(graphs):
            nf = np.asarray(_get(sample, "node_features"), dtype=np.float32)
            nf = nf.reshape(len(nf), -1)
            feats.append(nf)
            node_counts.append(len(nf))
            adj = _get(sample, "adjacency_lists")
            for t in range(self.num_edge_types):
                a = np.asarray(adj[t], dtype=np.int32).reshape(-1, 2)
                edges[t].append(a)
                edge_counts[t, g] = len(a)
        self.num_graphs = len(graphs)
        # host copies of the offset tables: sizes of a batch are known without a device round trip
        self.node_offsets_host = np.concatenate([[0], np.cumsum(node_counts, dtype=np.int64)]).astype(np.int64)
        self.edge_offsets_host = [np.concatenate([[0], np.cumsum(edge_counts[t])]).astype(np.int64)
                                  for t in range(self.num_edge_types)]
        dev = self.device
        self.node_features = torch.from_numpy(
            np.concatenate(feats, axis=0) if feats else np.zeros((0, 0), np.float32)).to(dev)
        self.node_offsets = torch.from_numpy(self.node_offsets_host).to(dev)
        self.edge_offsets = [torch.from_numpy(o).to(dev) for o in self.edge_offsets_host]
        self.edges = [torch.from_numpy(np.concatenate(e, axis=0) if e else np.zeros((0, 2), np.int32)).to(dev)
                      for e in edges]

    # ---- host logic: which graphs go into a batch (graph_dataset.py:164-188) ----------------------------------
    def iter_batch_graph_ids(self, max_nodes_per_batch: int, graph_order: Optional[Sequence[int]] = None
                             ) -> Iterator[np.ndarray]:
        return greedy_batches(np.diff(self.node_offsets_host), max_nodes_per_batch, graph_order)

    # ---- device: assemble the batch ---------------------------------------------------------------------------------
    def batch(self, graph_ids, with_node_features: bool = True) -> Dict[str, Any]:
        """batch_features of graph_dataset.py:226-246 as CUDA tensors: node_features, node_to_graph_map,
        num_graphs_in_batch, adjacency_list_{t}."""
        ids_host = np.asarray(graph_ids, dtype=np.int32).reshape(-1)
        if ids_host.size and (ids_host.min() < 0 or ids_host.max() >= self.num_graphs):
            raise IndexError("graph id out of range")
        dev = self.device
        Gb = int(ids_host.size)
        T = self.num_edge_types
        no, eo = self.node_offsets_host, self.edge_offsets_host
        Vb = int((no[ids_host + 1] - no[ids_host]).sum()) if Gb else 0
        Eb = [int((eo[t][ids_host + 1] - eo[t][ids_host]).sum()) if Gb else 0 for t in range(T)]
        ids = torch.from_numpy(ids_host).to(dev, non_blocking=True)
        n2g = torch.empty((Vb,), dtype=torch.int32, device=dev)
        rows = torch.empty((Vb,), dtype=torch.int32, device=dev) if with_node_features else None
        adj = [torch.empty((Eb[t], 2), dtype=torch.int32, device=dev) for t in range(T)]
        lib = _ffi.lib()
        ws = torch.empty((max(int(lib.tfgnn_b200_assemble_batch_workspace_bytes(T, Gb)), 8),), dtype=torch.uint8, device=dev)
        eoff_ptrs = (c_void_p * max(T, 1))(*[o.data_ptr() for o in self.edge_offsets])
        edge_ptrs = (c_void_p * max(T, 1))(*[e.data_ptr() if e.numel() else None for e in self.edges])
        out_ptrs = (c_void_p * max(T, 1))(*[a.data_ptr() if a.numel() else None for a in adj])
        Eb_c = (c_int64 * max(T, 1))(*Eb)
        _ffi.check(lib.tfgnn_b200_assemble_batch(
            self.node_offsets.data_ptr(), ctypes.cast(eoff_ptrs, _ffi._PP), ctypes.cast(edge_ptrs, _ffi._PP), T,
            self.num_graphs, ids.data_ptr() if Gb else None, Gb, Vb, Eb_c, n2g.data_ptr() if Vb else None,
            rows.data_ptr() if (rows is not None and Vb) else None, ctypes.cast(out_ptrs, _ffi._PP), ws.data_ptr(),
            stream_ptr()))
        features: Dict[str, Any] = {"node_to_graph_map": n2g, "num_graphs_in_batch": Gb}
        if with_node_features:
            F = int(self.node_features.shape[1]) if self.node_features.dim() == 2 else 0
            nf = torch.empty((Vb, F), dtype=torch.float32, device=dev)
            if Vb and F:
                _ffi.check(lib.tfgnn_b200_gather_rows(self.node_features.data_ptr(), int(self.node_features.shape[0]), F,
                                                      rows.data_ptr(), 1, Vb, nf.data_ptr(), stream_ptr()))
            features["node_features"] = nf
        for t in range(T):
            features[f"adjacency_list_{t}"] = adj[t]
        return features


def greedy_batches(node_counts: Sequence[int], max_nodes_per_batch: int,
                   graph_order: Optional[Sequence[int]] = None) -> Iterator[np.ndarray]:
    """The reference's batching rule (graph_dataset.py:164-188): graphs are taken in order and the batch under
    construction is emitted as soon as adding the next graph would exceed max_nodes_per_batch; a single over-sized
    graph still forms its own batch.  (When the very FIRST graph is over-sized the reference emits an empty batch and
    fails in np.concatenate; here empty batches are skipped.)  Yields int32 arrays of graph ids."""
    counts = np.asarray(node_counts, dtype=np.int64)
    order = np.arange(len(counts)) if graph_order is None else np.asarray(graph_order)
    cur: List[int] = []
    nodes = 0
    for g in order:
        n = int(counts[g])
        if nodes + n > max_nodes_per_batch:
            if cur:
                yield np.asarray(cur, dtype=np.int32)
            cur, nodes = [], 0
        cur.append(int(g))
        nodes += n
    if cur:
        yield np.asarray(cur, dtype=np.int32)


def _get(sample, name):
    return sample[name] if isinstance(sample, dict) else getattr(sample, name)
