"""The oracle against every golden vector the reference's tests hold for the hot path
(SURVEY.md §8c).  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import adjacency_oracle as ao
from oracle import message_passing_oracle as mo


def _load(golden_dir, name):
    with open(os.path.join(golden_dir, name)) as f:
        return json.load(f)


def test_pass_source_states_golden(golden_dir):
    """tf2_gnn/test/layers/test_message_passing.py:35-84 (assert_array_almost_equal, 6 decimals)."""
    g = _load(golden_dir, "message_passing_golden.json")
    assert len(g["pass_source_states"]) == 4
    for case in g["pass_source_states"]:
        params = mo.default_hyperparameters("pass_source_states")
        params.update(g["params"])
        params["hidden_dim"] = 3
        out = mo.message_passing_forward(
            "pass_source_states", params, {}, np.array(case["node_embeddings"], np.float32),
            [np.array(a, np.int32) for a in case["adjacency_lists"]])
        expected = np.array(case["aggregated_states"], np.float32)
        assert out.shape == expected.shape
        np.testing.assert_array_almost_equal(out, expected)


def test_in_degree_doctest_golden(golden_dir):
    """message_passing.py:238-249, exact values."""
    d = _load(golden_dir, "message_passing_golden.json")["in_degree_doctest"]
    got = mo.calculate_type_to_num_incoming_edges(d["num_nodes"], [np.array(a) for a in d["adjacency_lists"]])
    assert got.dtype == np.float32
    assert np.array_equal(got, np.array(d["type_to_num_incoming_edges"], np.float32))


def test_process_adjacency_lists_golden(golden_dir):
    """test/data/test_utils.py:50-138 + executed-reference fixtures: bit-exact."""
    g = _load(golden_dir, "process_adjacency_lists_golden.json")
    assert len(g["cases"]) >= 8
    for case in g["cases"]:
        i = case["input"]
        tied = ao.get_tied_edge_types(i["tie_fwd_bkwd_edges"], len(i["adjacency_lists"]))
        adj, cnt = ao.process_adjacency_lists(i["adjacency_lists"], i["num_nodes"], i["add_self_loop_edges"],
                                              tied, i["self_loop_edge_type"])
        assert len(adj) == len(case["adjacency_lists"])
        assert len(adj) == ao.compute_number_of_edge_types(tied, len(i["adjacency_lists"]),
                                                           i["add_self_loop_edges"])
        for got, exp in zip(adj, case["adjacency_lists"]):
            assert got.dtype == np.int32
            assert np.array_equal(got, np.array(exp, np.int32).reshape(-1, 2))
        assert np.array_equal(cnt, np.array(case["type_to_num_incoming_edges"]))


def test_in_degree_matches_process_adjacency_counts(golden_dir):
    """The two in-degree definitions of the reference agree (message_passing.py:230-263 vs
    data/utils.py:116-124)."""
    g = _load(golden_dir, "process_adjacency_lists_golden.json")
    for case in g["cases"]:
        n = case["input"]["num_nodes"]
        adj = [np.array(a, np.int32).reshape(-1, 2) for a in case["adjacency_lists"]]
        got = mo.calculate_type_to_num_incoming_edges(n, adj)
        assert np.array_equal(got, np.array(case["type_to_num_incoming_edges"], np.float32).reshape(len(adj), n))


@pytest.mark.parametrize("kind", ["rgcn", "ggnn", "gnn_edge_mlp", "gnn_film", "rgin", "rgat"])
def test_oracle_fp32_vs_fp64(kind):
    """Unpinned variants: the fp32 restatement stays within 1e-5*max|.| of its fp64 evaluation."""
    rng = np.random.default_rng(1)
    V, D, L = 200, 24, 3
    h = rng.uniform(-1, 1, (V, D)).astype(np.float32)
    adjs = [rng.integers(0, V, (int(rng.integers(0, 900)), 2)).astype(np.int32) for _ in range(L)]
    p = mo.default_hyperparameters(kind)
    p["hidden_dim"] = 24
    w = mo.make_weights(kind, p, D, L, rng)
    y32 = mo.message_passing_forward(kind, p, w, h, adjs)
    y64 = mo.message_passing_forward(kind, p, w, h, adjs, dtype=np.float64)
    assert y32.dtype == np.float32 and y32.shape == (V, 24)
    assert np.abs(y32 - y64).max() <= 1e-5 * np.abs(y64).max()


def test_aggregations_and_empty_segments():
    data = np.array([[1.0, -2.0], [3.0, 4.0], [5.0, -6.0]], np.float32)
    ids = np.array([2, 2, 0])
    assert np.array_equal(mo.unsorted_segment_sum(data, ids, 4), [[5, -6], [0, 0], [4, 2], [0, 0]])
    assert np.array_equal(mo.unsorted_segment_mean(data, ids, 4), [[5, -6], [0, 0], [2, 1], [0, 0]])
    np.testing.assert_allclose(mo.unsorted_segment_sqrt_n(data, ids, 4)[2], np.array([4, 2]) / np.sqrt(2), rtol=1e-6)
    mx = mo.unsorted_segment_max(data, ids, 4)
    assert np.array_equal(mx[2], [3, 4]) and mx[1, 0] == np.finfo(np.float32).min


def test_activation_table_errors():
    with pytest.raises(ValueError):
        mo.get_activation_function("linear")  # param_helpers.py:29-41 quirk
    with pytest.raises(ValueError):
        mo.get_aggregation_function("median")
    assert mo.get_activation_function(None) is None
    x = np.array([-1.0, 0.5], np.float32)
    np.testing.assert_allclose(mo.get_activation_function("ReLU")(x), [0, 0.5])
    np.testing.assert_allclose(mo.get_activation_function("leaky_relu")(x), [-0.2, 0.5], rtol=1e-6)


def test_assemble_batch_matches_executed_reference(golden_dir):
    """tests/golden/batch_assembly_golden.json: minibatches produced by EXECUTING the reference's own
    graph_dataset.py:161-246 (oracle/gen_golden.py) -> pins adjacency_oracle.assemble_batch bit-exactly."""
    import json
    import os
    import numpy as np
    from oracle import adjacency_oracle as ao
    with open(os.path.join(golden_dir, "batch_assembly_golden.json")) as f:
        cases = json.load(f)["cases"]
    assert len(cases) >= 4
    for c in cases:
        T = c["num_edge_types"]
        graphs = [dict(node_features=np.asarray(g["node_features"], np.float32),
                       adjacency_lists=[np.asarray(a, np.int32).reshape(-1, 2) for a in g["adjacency_lists"]])
                  for g in c["graphs"]]
        for b in c["batches"]:
            got = ao.assemble_batch([graphs[i] for i in b["graph_ids"]], T)
            assert got["num_graphs_in_batch"] == b["num_graphs_in_batch"]
            assert np.array_equal(got["node_to_graph_map"], np.asarray(b["node_to_graph_map"], np.int32))
            assert got["node_to_graph_map"].dtype == np.int32
            assert np.array_equal(got["node_features"], np.asarray(b["node_features"], np.float32))
            for t in range(T):
                expect = np.asarray(b["adjacency_lists"][t], np.int32).reshape(-1, 2)
                assert np.array_equal(got[f"adjacency_list_{t}"], expect)
                assert got[f"adjacency_list_{t}"].dtype == np.int32 and b["adjacency_dtypes"][t] == "int32"
