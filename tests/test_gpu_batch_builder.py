"""On-device batch builder against the pinned adjacency oracle and the reference's golden cases: bit-exact."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import adjacency_oracle as ao

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "process_adjacency_lists_golden.json")


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def test_process_adjacency_lists_matches_reference_golden_cases():
    """tests/golden/process_adjacency_lists_golden.json: outputs of the reference's own data/utils.py plus the 8
    expected results of test/data/test_utils.py:50-115."""
    _need_gpu()
    from tf2_gnn_b200.data import get_tied_edge_types, process_adjacency_lists
    with open(GOLDEN) as f:
        cases = json.load(f)["cases"]
    for c in cases:
        inp = c["input"]
        T = len(inp["adjacency_lists"])
        tied = get_tied_edge_types(inp["tie_fwd_bkwd_edges"], T)
        adjs = [np.asarray(a, np.int32).reshape(-1, 2) for a in inp["adjacency_lists"]]
        got, counts = process_adjacency_lists(adjs, inp["num_nodes"], inp["add_self_loop_edges"], tied,
                                              inp["self_loop_edge_type"])
        assert len(got) == len(c["adjacency_lists"])
        for g, e in zip(got, c["adjacency_lists"]):
            assert g.dtype == torch.int32
            assert np.array_equal(g.cpu().numpy(), np.asarray(e, np.int32).reshape(-1, 2))
        expect_counts = np.asarray(c["type_to_num_incoming_edges"], np.float64).reshape(len(got), inp["num_nodes"])
        assert np.array_equal(counts.cpu().numpy().astype(np.float64), expect_counts)


@pytest.mark.parametrize("V,sizes,tie,self_loops,self_type", [
    (1, (0,), False, True, 0),
    (50, (200, 0, 77), [1], True, 2),
    (50, (200, 0, 77), True, True, -1),
    (1000, (5000, 3000), False, False, 0),
    (300000, (1200000, 1, 400000), [0, 2], True, -3),
])
def test_process_adjacency_lists_matches_oracle(V, sizes, tie, self_loops, self_type):
    _need_gpu()
    from tf2_gnn_b200.data import get_tied_edge_types, process_adjacency_lists
    rng = np.random.default_rng(V + len(sizes))
    adjs = [rng.integers(0, V, size=(n, 2)).astype(np.int32) for n in sizes]
    tied = get_tied_edge_types(tie, len(sizes))
    expect, expect_counts = ao.process_adjacency_lists(adjs, V, self_loops, tied, self_type)
    got, counts = process_adjacency_lists([torch.from_numpy(a).cuda() for a in adjs], V, self_loops, tied, self_type)
    assert len(got) == len(expect)
    for g, e in zip(got, expect):
        assert np.array_equal(g.cpu().numpy(), e)
    assert np.array_equal(counts.cpu().numpy().astype(np.float64), expect_counts)


def test_process_adjacency_lists_rejects_bad_self_loop_slot():
    _need_gpu()
    from tf2_gnn_b200.data import process_adjacency_lists
    with pytest.raises(AssertionError):
        process_adjacency_lists([np.zeros((1, 2), np.int32)], 2, True, set(), 4)


def _random_graphs(rng, G, T, max_nodes, feature_dim):
    graphs = []
    for g in range(G):
        n = int(rng.integers(1, max_nodes + 1))
        adj = []
        for t in range(T):
            e = 0 if (t == 1 and g % 3 == 0) else int(rng.integers(0, 3 * n + 1))
            adj.append(rng.integers(0, n, size=(e, 2)).astype(np.int32))
        graphs.append({"node_features": rng.standard_normal((n, feature_dim)).astype(np.float32),
                       "adjacency_lists": adj})
    return graphs


@pytest.mark.parametrize("G,T,max_nodes,F", [(1, 1, 5, 3), (40, 3, 30, 15), (2500, 4, 29, 15), (6, 2, 3000, 50)])
def test_assemble_batch_matches_oracle(G, T, max_nodes, F):
    """graph_dataset.py:204-246: node ids offset by the running node count, constant node_to_graph_map blocks."""
    _need_gpu()
    from tf2_gnn_b200.data import DeviceGraphStore
    rng = np.random.default_rng(G * 7 + T)
    graphs = _random_graphs(rng, G, T, max_nodes, F)
    store = DeviceGraphStore(graphs, T)
    picks = [np.arange(G), rng.permutation(G)[: max(1, G // 2)], np.array([G - 1, 0, G - 1])]
    for ids in picks:
        expect = ao.assemble_batch([graphs[i] for i in ids], T)
        got = store.batch(ids)
        assert got["num_graphs_in_batch"] == expect["num_graphs_in_batch"]
        assert np.array_equal(got["node_to_graph_map"].cpu().numpy(), expect["node_to_graph_map"])
        assert np.array_equal(got["node_features"].cpu().numpy(), expect["node_features"])
        for t in range(T):
            a = got[f"adjacency_list_{t}"]
            assert a.dtype == torch.int32 and tuple(a.shape) == expect[f"adjacency_list_{t}"].shape
            assert np.array_equal(a.cpu().numpy(), expect[f"adjacency_list_{t}"])


def test_assemble_batch_matches_executed_reference_golden():
    """tests/golden/batch_assembly_golden.json: the reference's own graph_dataset.py:161-246, executed."""
    _need_gpu()
    from tf2_gnn_b200.data import DeviceGraphStore
    with open(os.path.join(os.path.dirname(GOLDEN), "batch_assembly_golden.json")) as f:
        cases = json.load(f)["cases"]
    for c in cases:
        T = c["num_edge_types"]
        graphs = [dict(node_features=np.asarray(g["node_features"], np.float32),
                       adjacency_lists=[np.asarray(a, np.int32).reshape(-1, 2) for a in g["adjacency_lists"]])
                  for g in c["graphs"]]
        store = DeviceGraphStore(graphs, T)
        ids_per_batch = [b.tolist() for b in store.iter_batch_graph_ids(c["max_nodes_per_batch"])]
        assert ids_per_batch == [b["graph_ids"] for b in c["batches"]]
        for b in c["batches"]:
            got = store.batch(b["graph_ids"])
            assert got["num_graphs_in_batch"] == b["num_graphs_in_batch"]
            assert np.array_equal(got["node_to_graph_map"].cpu().numpy(), np.asarray(b["node_to_graph_map"], np.int32))
            assert np.array_equal(got["node_features"].cpu().numpy(), np.asarray(b["node_features"], np.float32))
            for t in range(T):
                assert np.array_equal(got[f"adjacency_list_{t}"].cpu().numpy(),
                                      np.asarray(b["adjacency_lists"][t], np.int32).reshape(-1, 2))


def test_batches_from_the_store_drive_the_layer_like_host_built_ones():
    """The whole device-side data path: packed store -> batch -> process_adjacency_lists -> RGCN layer, equal to the
    same layer on the oracle-built batch."""
    _need_gpu()
    from tf2_gnn_b200.data import DeviceGraphStore, process_adjacency_lists
    from tf2_gnn_b200.layers.message_passing import MessagePassingInput, get_message_passing_class
    rng = np.random.default_rng(11)
    T, F = 2, 32
    graphs = _random_graphs(rng, 60, T, 25, F)
    store = DeviceGraphStore(graphs, T)
    ids = next(iter(store.iter_batch_graph_ids(max_nodes_per_batch=400)))
    b = store.batch(ids)
    V = int(b["node_to_graph_map"].shape[0])
    adjs, _ = process_adjacency_lists([b[f"adjacency_list_{t}"] for t in range(T)], V, True, {0}, 0)
    hb = ao.assemble_batch([graphs[i] for i in ids], T)
    adjs_ref, _ = ao.process_adjacency_lists([hb[f"adjacency_list_{t}"] for t in range(T)], V, True, {0}, 0)
    cls = get_message_passing_class("rgcn")
    params = cls.get_default_hyperparameters()
    params["hidden_dim"] = 32
    layer = cls(params)
    torch.manual_seed(0)
    layer.build(MessagePassingInput((None, F), tuple((None, 2) for _ in adjs)))
    out = layer(MessagePassingInput(b["node_features"], tuple(adjs)))
    out_ref = layer(MessagePassingInput(torch.from_numpy(hb["node_features"]).cuda(),
                                        tuple(torch.from_numpy(a).cuda() for a in adjs_ref)))
    assert torch.equal(out, out_ref)
