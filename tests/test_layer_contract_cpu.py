"""The plug-in contract of SURVEY.md §8(b), checked without a GPU (building a layer only creates its weights):
 * number / shape / names of the trainable variables, mirroring the reference's own tests
   (tf2_gnn/test/layers/test_RGCN.py:8-65, test_RGAT.py:8-64);
 * class registry keyed by lower-cased class name (message_passing/__init__.py:10-14, message_passing.py:221-227);
 * default hyper-parameters of every message-passing class (message_passing.py:41-48, gnn_edge_mlp.py:46-55,
   rgcn.py:50-59, ggnn.py:47-56, rgat.py:53-60, gnn_film.py:49-59, rgin.py:61-70)."""
import pytest

from tf2_gnn_b200.layers import GNN, MessagePassingInput, get_message_passing_class
from tf2_gnn_b200.layers.message_passing import GGNN, GNN_Edge_MLP, GNN_FiLM, RGAT, RGCN, RGIN

shape_test_data = [
    ((None, 3), tuple((None, 2) for _ in range(3)), 5),
    ((None, 1), tuple((None, 2) for _ in range(1)), 1),
    ((None, 7), tuple((None, 2) for _ in range(14)), 7),
]


@pytest.mark.parametrize("node_embedding_shape,adjacency_list_shapes,hidden_dim", shape_test_data)
@pytest.mark.parametrize("use_target", [False, True])
def test_rgcn_layer_has_expected_trainable_variables(node_embedding_shape, adjacency_list_shapes, hidden_dim,
                                                     use_target):
    """test_RGCN.py:15-65: one bias-free dense kernel per edge type, [D or 2D, hidden_dim]."""
    params = RGCN.get_default_hyperparameters()
    params["hidden_dim"] = hidden_dim
    params["use_target_state_as_input"] = use_target
    layer = RGCN(params)
    layer.build(MessagePassingInput(node_embeddings=node_embedding_shape, adjacency_lists=adjacency_list_shapes))
    trainable_vars, all_vars = layer.trainable_variables, layer.variables
    assert len(trainable_vars) == len(adjacency_list_shapes)
    assert len(all_vars) == len(trainable_vars)
    in_dim = (2 if use_target else 1) * node_embedding_shape[-1]
    for v in trainable_vars:
        assert tuple(v.shape.as_list()) == (in_dim, hidden_dim)
        assert "bias" not in v.name


rgat_shape_test_data = [
    ((None, 3), tuple((None, 2) for _ in range(3)), 16, 8),
    ((None, 1), tuple((None, 2) for _ in range(1)), 2, 1),
    ((None, 7), tuple((None, 2) for _ in range(14)), 64, 4),
]


@pytest.mark.parametrize("node_embedding_shape,adjacency_list_shapes,hidden_dim,num_heads", rgat_shape_test_data)
def test_rgat_layer_has_expected_trainable_variables(node_embedding_shape, adjacency_list_shapes, hidden_dim,
                                                     num_heads):
    """test_RGAT.py:31-64: one dense kernel [D, H] and one attention weight [K, 2H/K] per edge type."""
    params = RGAT.get_default_hyperparameters()
    params["hidden_dim"] = hidden_dim
    params["num_heads"] = num_heads
    layer = RGAT(params)
    layer.build(MessagePassingInput(node_embeddings=node_embedding_shape, adjacency_lists=adjacency_list_shapes))
    trainable_vars, all_vars = layer.trainable_variables, layer.variables
    assert len(trainable_vars) == 2 * len(adjacency_list_shapes)
    assert len(all_vars) == len(trainable_vars)
    for v in trainable_vars:
        if "kernel" in v.name:
            assert tuple(v.shape.as_list()) == (node_embedding_shape[-1], hidden_dim)
        elif "attention" in v.name:
            assert tuple(v.shape.as_list()) == (num_heads, 2 * hidden_dim // num_heads)
        else:
            raise AssertionError(f"unexpected variable {v.name}")


def test_ggnn_gru_cell_variables():
    """ggnn.py:62-66: Keras GRUCell(units=H): kernel [D,3H], recurrent_kernel [H,3H], bias [2,3H] (reset_after)."""
    params = GGNN.get_default_hyperparameters()
    params["hidden_dim"] = 12
    layer = GGNN(params)
    layer.build(MessagePassingInput((None, 12), tuple((None, 2) for _ in range(2))))
    gru = [tuple(v.shape.as_list()) for v in layer.variables if "gru_cell" in v.name]
    assert sorted(gru) == sorted([(12, 36), (12, 36), (2, 36)])
    assert len(layer.variables) == 3 + 2   # + one message kernel per edge type


def test_registry_is_keyed_by_lower_cased_class_name():
    for name, cls in [("rgcn", RGCN), ("RGCN", RGCN), ("rgat", RGAT), ("rgin", RGIN), ("ggnn", GGNN),
                      ("gnn_edge_mlp", GNN_Edge_MLP), ("GNN_Edge_MLP", GNN_Edge_MLP), ("gnn_film", GNN_FiLM)]:
        assert get_message_passing_class(name) is cls
    with pytest.raises(ValueError):
        get_message_passing_class("no_such_layer")


BASE = {"aggregation_function": "sum", "message_activation_function": "relu",
        "message_activation_before_aggregation": False, "hidden_dim": 7}


@pytest.mark.parametrize("cls,extra", [
    (GNN_Edge_MLP, {"use_target_state_as_input": True, "normalize_by_num_incoming": False,
                    "num_edge_MLP_hidden_layers": 1}),
    (RGCN, {"use_target_state_as_input": False, "normalize_by_num_incoming": True, "num_edge_MLP_hidden_layers": 0}),
    (GGNN, {"use_target_state_as_input": False, "normalize_by_num_incoming": True, "num_edge_MLP_hidden_layers": 0}),
    (RGAT, {"num_heads": 3}),
    (GNN_FiLM, {"film_parameter_MLP_hidden_layers": []}),
    (RGIN, {"use_target_state_as_input": False, "num_edge_MLP_hidden_layers": 1, "num_aggr_MLP_hidden_layers": None}),
])
def test_default_hyperparameters_match_the_reference(cls, extra):
    got = cls.get_default_hyperparameters()
    for k, v in {**BASE, **extra}.items():
        assert k in got, f"{cls.__name__} lacks hyper-parameter {k}"
        assert got[k] == v, f"{cls.__name__}.{k}: {got[k]!r} != {v!r}"


def test_gnn_default_hyperparameters_carry_the_message_passing_ones():
    """gnn.py:29-79: GNN.get_default_hyperparameters(mp_style) merges the layer's defaults under its own keys."""
    p = GNN.get_default_hyperparameters("rgcn")
    assert p["message_calculation_class"] == "rgcn"
    for k in ("num_layers", "hidden_dim", "dense_every_num_layers", "residual_every_num_layers",
              "use_inter_layer_layernorm", "layer_input_dropout_rate", "global_exchange_mode",
              "global_exchange_every_num_layers", "initial_node_representation_activation"):
        assert k in p, k
    assert p["normalize_by_num_incoming"] is True and p["num_edge_MLP_hidden_layers"] == 0
