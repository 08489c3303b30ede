"""CPU sanity of the readout / exchange oracle (numpy) and of the host-side layer construction (no compute calls)."""
import numpy as np
import pytest

from oracle import message_passing_oracle as mo


def test_softmax_weights_sum_to_one_per_graph_and_head():
    rng = np.random.default_rng(0)
    V, G, K = 200, 9, 3
    ids = np.sort(rng.integers(0, G, V))
    s = rng.normal(size=(V, K))
    w = np.stack([mo.unsorted_segment_softmax(s[:, k], ids, G) for k in range(K)], 1)
    sums = mo.unsorted_segment_sum(w, ids, G)
    present = np.bincount(ids, minlength=G) > 0
    assert np.allclose(sums[present], 1.0) and np.all(sums[~present] == 0)


def test_readout_none_and_average_reduce_to_segment_ops():
    rng = np.random.default_rng(1)
    V, G, D, GD = 120, 5, 8, 6
    x = rng.uniform(-1, 1, (V, D)).astype(np.float32)
    ids = np.sort(rng.integers(0, G, V))
    w = {"transformation_mlp": [mo.glorot_uniform(rng, (D, 7)), mo.glorot_uniform(rng, (7, GD))]}
    t = np.maximum(mo.mlp_forward(x.astype(np.float64), [m.astype(np.float64) for m in w["transformation_mlp"]]), 0)
    none = mo.weighted_sum_graph_representation(x, ids, G, w, GD, 2, "none", dtype=np.float64)
    avg = mo.weighted_sum_graph_representation(x, ids, G, w, GD, 2, "average", dtype=np.float64)
    assert np.allclose(none, mo.unsorted_segment_sum(t, ids, G))
    assert np.allclose(avg, mo.unsorted_segment_mean(t, ids, G))


def test_exchange_modes_shapes_and_mean_identity():
    rng = np.random.default_rng(2)
    V, G, H = 90, 4, 8
    x = rng.uniform(-1, 1, (V, H)).astype(np.float32)
    ids = np.sort(rng.integers(0, G, V))
    for mode in ("mean", "gru", "mlp"):
        w = mo.make_exchange_weights(mode, H, 4, rng)
        out = mo.graph_global_exchange(mode, x, ids, G, w, H, 4, "softmax")
        assert out.shape == (V, H) and np.isfinite(out).all()
    w = mo.make_exchange_weights("mean", H, 4, rng)
    g = mo.weighted_sum_graph_representation(x, ids, G, w, H, 4, "softmax")
    assert np.allclose(mo.graph_global_exchange("mean", x, ids, G, w, H, 4, "softmax"), (x + g[ids]) / 2)
    with pytest.raises(ValueError):
        mo.graph_global_exchange("sum", x, ids, G, w, H, 4)


def test_default_gnn_builds_its_exchange_layers_with_reference_shapes():
    """gnn.py:172-200 + graph_global_exchange.py:46-58: default hypers -> one GRU exchange at layer 2."""
    from tf2_gnn_b200.layers import GNN, GNNInput
    params = GNN.get_default_hyperparameters()
    gnn = GNN(params)
    gnn.build(GNNInput((None, 10), ((None, 2), (None, 2)), (None,), ()))
    assert sorted(gnn._global_exchange_layers) == ["2"]
    ex = gnn._global_exchange_layers["2"]
    H = params["hidden_dim"]
    shapes = {v.name.split("Global_Exchange/")[1]: tuple(v.shape) for v in ex.variables}
    assert shapes["GraphGlobalGRUExchange/gru_cell/kernel:0"] == (H, 3 * H)
    assert shapes["GraphGlobalGRUExchange/gru_cell/bias:0"] == (2, 3 * H)
    assert shapes["GraphGlobalGRUExchange/WeightedSumGraphRepresentation/ScoringMLP/dense_0/kernel:0"] == (H, H)
    assert shapes["GraphGlobalGRUExchange/WeightedSumGraphRepresentation/ScoringMLP/dense_1/kernel:0"] == (H, 4)
    assert shapes["GraphGlobalGRUExchange/WeightedSumGraphRepresentation/TransformationMLP/dense_0/kernel:0"] == (H, 128)
    assert shapes["GraphGlobalGRUExchange/WeightedSumGraphRepresentation/TransformationMLP/dense_1/kernel:0"] == (128, H)
    with pytest.raises(ValueError):
        GNN(dict(params, global_exchange_mode="sum"))
