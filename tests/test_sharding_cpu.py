"""Multi-GPU host logic on CPU: partitioning is bit-exact and the N>1 data flow (world_size 2, gloo)
reproduces the single-process result.  The per-rank layer compute is the oracle here (no GPU in CI);
the GPU test of the same flow is tests/test_gpu_parity.py::test_target_range_shards_match_full."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import message_passing_oracle as mo
from tf2_gnn_b200 import sharding


def make_disjoint_batch(rng, num_graphs=7, L=3):
    feats, n2g, adj = [], [], [[] for _ in range(L)]
    off = 0
    for g in range(num_graphs):
        n = int(rng.integers(5, 40))
        feats.append(rng.uniform(-1, 1, (n, 16)).astype(np.float32))
        n2g.append(np.full(n, g, np.int32))
        for l in range(L):
            e = rng.integers(0, n, size=(int(rng.integers(0, 4 * n)), 2)).astype(np.int32) + off
            adj[l].append(e)
        off += n
    return np.concatenate(feats), [np.concatenate(a) for a in adj], np.concatenate(n2g)


def test_partition_by_graph_cuts_only_at_graph_boundaries():
    rng = np.random.default_rng(0)
    feats, adjs, n2g = make_disjoint_batch(rng, 11)
    for world in (1, 2, 3, 8, 16):
        bounds = sharding.partition_by_graph(n2g, adjs, world)
        assert len(bounds) == world and bounds[0][0] == 0 and bounds[-1][1] == len(n2g)
        for (lo, hi), (lo2, _) in zip(bounds[:-1], bounds[1:]):
            assert hi == lo2
        for lo, hi in bounds:
            if 0 < lo < len(n2g):
                assert n2g[lo] != n2g[lo - 1]
        total_edges = 0
        for r in range(world):
            sh = sharding.shard_disjoint_batch(feats, adjs, n2g, bounds, r)
            lo, hi = sh["node_range"]
            for a, full in zip(sh["adjacency_lists"], adjs):
                assert a.dtype == np.int32
                keep = (full[:, 1] >= lo) & (full[:, 1] < hi)
                assert np.array_equal(a + lo, full[keep])          # order preserved, exact re-basing
                total_edges += len(a)
                assert len(a) == 0 or (a.min() >= 0 and a.max() < hi - lo)
        assert total_edges == sum(len(a) for a in adjs)


def test_shard_rejects_cross_graph_edges():
    n2g = np.array([0, 0, 1, 1], np.int32)
    adjs = [np.array([[0, 3]], np.int32)]
    with pytest.raises(ValueError):
        sharding.shard_disjoint_batch(np.zeros((4, 2), np.float32), adjs, n2g, [(0, 2), (2, 4)], 0)


def test_partition_target_range_covers_all_rows():
    for V, world in [(10, 3), (1000, 8), (5, 8), (0, 2)]:
        b = sharding.partition_target_range(V, world)
        assert b[0][0] == 0 and b[-1][1] == V and all(x[1] == y[0] for x, y in zip(b[:-1], b[1:]))
    deg = np.zeros(100, np.int64)
    deg[:10] = 1000
    b = sharding.partition_target_range(100, 4, deg)
    assert b[0][1] <= 10 and b[-1][1] == 100


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(123)
        p = mo.default_hyperparameters("rgcn")
        p["hidden_dim"] = 16
        # ---- case 1: disjoint graphs, no collective in the forward pass -------------------------
        feats, adjs, n2g = make_disjoint_batch(rng, 9)
        w = mo.make_weights("rgcn", p, 16, 3, rng)
        bounds = sharding.partition_by_graph(n2g, adjs, world)
        sh = sharding.shard_disjoint_batch(feats, adjs, n2g, bounds, rank)
        local = mo.message_passing_forward("rgcn", p, w, sh["node_features"], sh["adjacency_lists"])
        full1 = sharding.all_gather_node_states(torch.from_numpy(local), bounds).numpy()
        # ---- case 2: one graph, target-range shards + per-layer all-gather, 2 layers -------------
        V = 101
        h = rng.uniform(-1, 1, (V, 16)).astype(np.float32)
        adjs2 = [rng.integers(0, V, (300, 2)).astype(np.int32) for _ in range(3)]
        w2 = [mo.make_weights("rgcn", p, 16, 3, rng) for _ in range(2)]
        tb = sharding.partition_target_range(V, world)
        lo, hi = tb[rank]
        mine = sharding.filter_edges_by_target(adjs2, lo, hi)
        h_local = h[lo:hi]
        for layer in range(2):
            h_full = sharding.all_gather_node_states(torch.from_numpy(np.ascontiguousarray(h_local)), tb).numpy()
            out_full = mo.message_passing_forward("rgcn", p, w2[layer], h_full, mine)   # rows outside [lo,hi) unused
            h_local = out_full[lo:hi]
        full2 = sharding.all_gather_node_states(torch.from_numpy(np.ascontiguousarray(h_local)), tb).numpy()
        if rank == 0:
            ref1 = mo.message_passing_forward("rgcn", p, w, feats, adjs)
            ref2 = h
            for layer in range(2):
                ref2 = mo.message_passing_forward("rgcn", p, w2[layer], ref2, adjs2)
            np.save(os.path.join(tmp, "ok.npy"), np.array([
                np.abs(full1 - ref1).max() <= 1e-6 * max(np.abs(ref1).max(), 1),
                np.abs(full2 - ref2).max() <= 1e-6 * max(np.abs(ref2).max(), 1)]))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_matches_single_process(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ok = np.load(os.path.join(str(tmp_path), "ok.npy"))
    assert ok.all()
