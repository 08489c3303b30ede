"""Multi-GPU partitioning of the message-passing path.  One process per GPU.

Two cases, both with host-side index bookkeeping only (numpy; bit-exact, tested on CPU with gloo):

1. A batch of DISJOINT graphs (every dataset of the reference: graph_dataset.py:218-222 offsets each
   graph's node ids, so edges never cross graphs and node_to_graph_map is non-decreasing).  Cut the
   batch at graph boundaries into `world_size` contiguous node ranges balanced by edge count and
   re-base the indices per shard.  The forward pass needs NO collective.

2. ONE graph larger than a shard: 1-D partition by target-node range.  Rank g owns output rows
   [lo_g, hi_g) and every edge whose target falls there (ids stay global); it needs h[src] for
   arbitrary sources, i.e. one all-gather of the node-state shards per layer
   (tfgnn_b200_prepare_sharded + torch.distributed.all_gather_into_tensor over NCCL/NVLink).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


# ------------------------------------------------------------------------------------------------
# 1. disjoint-graph batches: cut at graph boundaries
# ------------------------------------------------------------------------------------------------
def partition_by_graph(node_to_graph_map: np.ndarray, adjacency_lists: Sequence[np.ndarray],
                       world_size: int) -> List[Tuple[int, int]]:
    """Contiguous node ranges [(lo, hi)] per rank, cut only at graph boundaries, balanced by the
    number of edges (the HBM traffic of the layer is proportional to it)."""
    n2g = np.asarray(node_to_graph_map)
    V = int(n2g.shape[0])
    if V == 0:
        return [(0, 0)] * world_size
    if np.any(np.diff(n2g) < 0):
        raise ValueError("node_to_graph_map must be non-decreasing (graph_dataset.py:211-217)")
    num_graphs = int(n2g[-1]) + 1
    graph_start = np.searchsorted(n2g, np.arange(num_graphs), side="left")          # first node of each graph
    edges_per_graph = np.zeros(num_graphs, dtype=np.int64)
    for adj in adjacency_lists:
        adj = np.asarray(adj).reshape(-1, 2)
        if len(adj):
            edges_per_graph += np.bincount(n2g[adj[:, 1]], minlength=num_graphs)
    load = edges_per_graph + np.diff(np.append(graph_start, V))                    # edges + nodes per graph
    cum = np.cumsum(load)
    total = int(cum[-1])
    bounds, g_lo = [], 0
    for r in range(world_size):
        if r == world_size - 1:
            g_hi = num_graphs
        else:
            target = total * (r + 1) / world_size
            g_hi = int(np.searchsorted(cum, target, side="left")) + 1
            g_hi = min(max(g_hi, g_lo), num_graphs)
        lo = int(graph_start[g_lo]) if g_lo < num_graphs else V
        hi = int(graph_start[g_hi]) if g_hi < num_graphs else V
        bounds.append((lo, hi))
        g_lo = g_hi
    return bounds


def shard_disjoint_batch(node_features: np.ndarray, adjacency_lists: Sequence[np.ndarray],
                         node_to_graph_map: np.ndarray, bounds: Sequence[Tuple[int, int]], rank: int
                         ) -> Dict[str, object]:
    """The rank's sub-batch with indices re-based to its node range.  Raises if an edge crosses the cut
    (it cannot for batches built by graph_dataset.py)."""
    lo, hi = bounds[rank]
    out_adj = []
    for adj in adjacency_lists:
        adj = np.asarray(adj, dtype=np.int32).reshape(-1, 2)
        tgt_in = (adj[:, 1] >= lo) & (adj[:, 1] < hi)
        src_in = (adj[:, 0] >= lo) & (adj[:, 0] < hi)
        if np.any(tgt_in != src_in):
            raise ValueError("an edge crosses a graph boundary: not a batch of disjoint graphs")
        out_adj.append(np.ascontiguousarray(adj[tgt_in] - np.int32(lo)))
    n2g = np.asarray(node_to_graph_map)[lo:hi]
    first_graph = int(n2g[0]) if hi > lo else 0
    return {
        "node_features": np.asarray(node_features)[lo:hi],
        "adjacency_lists": out_adj,
        "node_to_graph_map": (n2g - first_graph).astype(np.int32),
        "num_graphs": int(n2g[-1]) - first_graph + 1 if hi > lo else 0,
        "node_range": (lo, hi),
    }


# ------------------------------------------------------------------------------------------------
# 2. one large graph: 1-D partition by target range
# ------------------------------------------------------------------------------------------------
def partition_target_range(num_nodes: int, world_size: int, in_degree: Optional[np.ndarray] = None
                           ) -> List[Tuple[int, int]]:
    """Contiguous target ranges per rank; balanced by in-degree (edges) when given, else by rows."""
    if in_degree is None:
        cuts = [(num_nodes * r) // world_size for r in range(world_size + 1)]
    else:
        cum = np.cumsum(np.asarray(in_degree, dtype=np.int64) + 1)
        total = int(cum[-1]) if num_nodes else 0
        cuts = [0] + [int(np.searchsorted(cum, total * r / world_size, side="left")) for r in range(1, world_size)]
        cuts.append(num_nodes)
        cuts = [min(max(c, 0), num_nodes) for c in cuts]
        cuts = list(np.maximum.accumulate(cuts))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world_size)]


def filter_edges_by_target(adjacency_lists: Sequence[np.ndarray], lo: int, hi: int) -> List[np.ndarray]:
    """Edges whose target lies in [lo, hi); ids stay GLOBAL (tfgnn_b200_prepare_sharded re-bases the
    targets itself).  Optional: the library also accepts the unfiltered lists."""
    out = []
    for adj in adjacency_lists:
        adj = np.asarray(adj, dtype=np.int32).reshape(-1, 2)
        keep = (adj[:, 1] >= lo) & (adj[:, 1] < hi)
        out.append(np.ascontiguousarray(adj[keep]))
    return out


def padded_shard_rows(bounds: Sequence[Tuple[int, int]]) -> int:
    """Rows per rank for an equal-size all_gather_into_tensor (ranges are padded to the largest)."""
    return max((hi - lo) for lo, hi in bounds) if bounds else 0


def assemble_gathered(gathered: np.ndarray, bounds: Sequence[Tuple[int, int]]) -> np.ndarray:
    """[world * padded_rows, D] (concatenated padded shards) -> [V, D]."""
    rows = padded_shard_rows(bounds)
    parts = [gathered[r * rows: r * rows + (hi - lo)] for r, (lo, hi) in enumerate(bounds)]
    return np.concatenate(parts, axis=0) if parts else gathered[:0]


def all_gather_node_states(h_local, bounds: Sequence[Tuple[int, int]], group=None):
    """torch tensors on any device/backend: all-gather the per-rank row ranges into the full [V, D] table."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rows = padded_shard_rows(bounds)
    D = int(h_local.shape[1])
    send = h_local
    if int(h_local.shape[0]) != rows:
        send = torch.zeros((rows, D), dtype=h_local.dtype, device=h_local.device)
        send[: h_local.shape[0]] = h_local
    recv = torch.empty((world * rows, D), dtype=h_local.dtype, device=h_local.device)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    if all((hi - lo) == rows for lo, hi in bounds):
        return recv
    return torch.cat([recv[r * rows: r * rows + (hi - lo)] for r, (lo, hi) in enumerate(bounds)], dim=0)


# ------------------------------------------------------------------------------------------------
# 2b. the all-gather fused into the layer kernel: node-state tables in peer-mapped (symmetric) memory
# ------------------------------------------------------------------------------------------------
class PeerNodeTables:
    """Two [num_nodes, H] float32 node-state tables per rank, allocated in symmetric memory
    (torch.distributed._symmetric_memory: CUDA VMM allocations every rank of the group maps over NVLink), so that the
    epilogue of the fused layer kernel can store its output tiles straight into every rank's copy
    (GNN_Edge_MLP.call_allgather / tfgnn_b200_rgcn_fwd_allgather).  Layer k reads `table(k)` and writes `table(k + 1)` on
    all ranks; `barrier(k + 1)` is the rank synchronisation between layers.  Collective constructor."""

    def __init__(self, num_nodes: int, hidden_dim: int, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group = group or dist.group.WORLD
        dev = torch.device("cuda", torch.cuda.current_device())
        self._tables, self._handles = [], []
        for _ in range(2):
            t = symm_mem.empty((int(num_nodes), int(hidden_dim)), dtype=torch.float32, device=dev)
            self._handles.append(symm_mem.rendezvous(t, group))
            self._tables.append(t)
        self.rank = int(self._handles[0].rank)
        self.world_size = int(self._handles[0].world_size)

    def table(self, k: int):
        return self._tables[k % 2]

    def replica_ptrs(self, k: int):
        """Device addresses of table k on every rank, as mapped into this process (index = rank)."""
        return [int(p) for p in self._handles[k % 2].buffer_ptrs]

    def multicast_ptr(self, k: int) -> int:
        """Multicast (NVSwitch / NVLS) address of table k: one store to it lands in every rank's table; 0 if unsupported."""
        h = self._handles[k % 2]
        try:
            return int(h.multicast_ptr) if h.has_multicast_support(h.device.type, h.device.index) else 0
        except TypeError:
            try:
                return int(h.multicast_ptr or 0)
            except Exception:
                return 0
        except Exception:
            return 0

    def barrier(self, k: int) -> None:
        """All ranks have finished writing table k (enqueued on the current stream: a signal exchange in device memory)."""
        self._handles[k % 2].barrier(channel=0)
