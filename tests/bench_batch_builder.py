"""Measurement for the on-device batch builder: QM9-shaped (cfg4) and PPI-shaped (cfg1) minibatches.

Times `DeviceGraphStore.batch` + `process_adjacency_lists` on the GPU (CUDA events, inputs resident in HBM) next to
the reference's host path restated in numpy (oracle/adjacency_oracle.py: assemble_batch + process_adjacency_lists,
pinned bit-exactly against the reference's own code), and checks the two results against each other.
Lives under tests/ because it uses the oracle as checker and CPU baseline (only tests/, smoke() and bench.py may).
Prints one JSON line per workload.  Run on a GPU box: python tests/bench_batch_builder.py
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import adjacency_oracle as ao  # noqa: E402  (checker / CPU baseline only)
from tf2_gnn_b200.data import DeviceGraphStore, process_adjacency_lists  # noqa: E402


def qm9_like(rng, num_graphs):
    graphs = []
    for _ in range(num_graphs):
        n = int(rng.integers(9, 30))
        parents = [int(rng.integers(0, i)) for i in range(1, n)]
        bonds = np.array([[p, i + 1] for i, p in enumerate(parents)], np.int32)
        extra = rng.integers(0, n, size=(max(1, n // 20), 2)).astype(np.int32)
        allb = np.concatenate([bonds, extra])
        types = rng.choice(4, size=len(allb), p=(0.85, 0.1, 0.01, 0.04))
        graphs.append({"node_features": np.eye(15, dtype=np.float32)[rng.integers(0, 15, n)],
                       "adjacency_lists": [allb[types == t] for t in range(4)]})
    return graphs, dict(tied={0, 1, 2, 3}, self_loops=True)


def ppi_like(rng, num_graphs):
    graphs = []
    for _ in range(num_graphs):
        n = int(rng.integers(1800, 3500))
        e = int(14.4 * n)
        graphs.append({"node_features": rng.standard_normal((n, 50)).astype(np.float32),
                       "adjacency_lists": [rng.integers(0, n, size=(e, 2)).astype(np.int32)]})
    return graphs, dict(tied=set(), self_loops=True)


def run(name, graphs, opts, T, reps=20):
    store = DeviceGraphStore(graphs, T)
    ids = np.arange(len(graphs), dtype=np.int32)

    def device_path():
        b = store.batch(ids)
        V = int(b["node_to_graph_map"].shape[0])
        adjs, counts = process_adjacency_lists([b[f"adjacency_list_{t}"] for t in range(T)], V, opts["self_loops"],
                                               opts["tied"], 0)
        return b, adjs, counts

    b, adjs, counts = device_path()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    ev0.record()
    for _ in range(reps):
        device_path()
    ev1.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t_wall) * 1e3 / reps
    dev_ms = ev0.elapsed_time(ev1) / reps

    t0 = time.perf_counter()
    hb = ao.assemble_batch(graphs, T)
    V = len(hb["node_to_graph_map"])
    ref_adjs, ref_counts = ao.process_adjacency_lists([hb[f"adjacency_list_{t}"] for t in range(T)], V,
                                                      opts["self_loops"], opts["tied"], 0)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    ok = (np.array_equal(b["node_to_graph_map"].cpu().numpy(), hb["node_to_graph_map"])
          and all(np.array_equal(a.cpu().numpy(), r) for a, r in zip(adjs, ref_adjs))
          and np.array_equal(counts.cpu().numpy().astype(np.float64), ref_counts))
    edges_in = sum(int(hb[f"adjacency_list_{t}"].shape[0]) for t in range(T))
    edges_out = sum(int(a.shape[0]) for a in adjs)
    F = graphs[0]["node_features"].shape[1]
    # algorithmic bytes: packed edges read + batch edges written (8 B each), batch edges read + processed edges written,
    # node map + features written/read, in-degree table written
    alg = 16 * edges_in + 8 * (edges_in + edges_out) + V * (4 + 8 * F) + 4 * len(adjs) * V
    print(json.dumps({
        "workload": name, "graphs": len(graphs), "nodes": V, "edges_in": edges_in, "edges_processed": edges_out,
        "edge_types_out": len(adjs), "bit_exact_vs_oracle": bool(ok), "device_ms": dev_ms, "wall_ms_per_batch": wall_ms,
        "cpu_numpy_ms": cpu_ms, "algorithmic_bytes": alg, "achieved_GBps": alg / (dev_ms * 1e-3) / 1e9,
        "note": "device_ms = CUDA events over assemble_batch + gather_rows + process_adjacency (launch-bound at these "
                "sizes); cpu = numpy restatement of the reference's host loops on this box, 1 thread"}), flush=True)


def main():
    rng = np.random.default_rng(0)
    g, o = qm9_like(rng, 27800)
    run("cfg4 QM9-shaped: 27.8k molecules, 4 bond types tied + self loops", g, o, 4)
    g, o = ppi_like(rng, 3)
    run("cfg1 PPI-shaped: 3 graphs, 1 link type + backward + self loops", g, o, 1)


if __name__ == "__main__":
    main()
