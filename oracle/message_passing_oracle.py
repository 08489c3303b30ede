"""CPU oracle for the tf2_gnn message-passing hot path.  TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement, op for op, of the reference's
gather -> per-edge-type message -> unsorted_segment_* loop.  It is the checker that the
CUDA path is compared against; nothing in the product package (tf2_gnn_b200/) may import
it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg use it.

Pinning status (see DESIGN.md "Oracle"):
  * PINNED against the reference's own golden vectors:
      - gather -> identity message -> unsorted_segment_sum -> relu
        (tf2_gnn/test/layers/test_message_passing.py:35-71, 4 cases),
      - calculate_type_to_num_incoming_edges doctest (message_passing.py:238-249),
    see tests/golden/message_passing_golden.json and tests/test_oracle_golden.py.
  * PARITY UNPINNED (the reference holds no numeric test for them, and neither TensorFlow
    nor dpu_utils exists in this image, so the reference cannot be run): the Dense/MLP
    message transform, 1/(c+1e-7) scaling, RGAT scores + segment softmax, FiLM
    modulation, GGNN GRUCell, RGIN aggregation MLP, mean/max/sqrt_n, gelu/elu/selu.
    For these the restatement below follows the cited reference lines and the published
    semantics of the external ops (TensorFlow >=2.0 `tf.math.unsorted_segment_*`,
    `tf.keras.layers.GRUCell` with TF2 defaults, `dpu_utils.tf2utils.MLP` and
    `unsorted_segment_log_softmax`, dpu-utils>=0.2.7 — both unpinned in
    /root/reference/setup.py:22-29); it is cross-checked in float64 (dtype argument).

All file:line citations are relative to /root/reference/.
"""
from __future__ import annotations

import math
from typing import Any, Dict, Optional, Sequence, Tuple

import numpy as np

SMALL_NUMBER = 1e-7  # tf2_gnn/utils/constants.py:2
LEAKY_RELU_ALPHA = 0.2  # tf.nn.leaky_relu default alpha [external TF]
SELU_ALPHA = 1.6732632423543772
SELU_SCALE = 1.0507009873554805


# --------------------------------------------------------------------------------------
# Name -> op maps (tf2_gnn/utils/param_helpers.py:7-42, tf2_gnn/utils/activation.py:7-14)
# --------------------------------------------------------------------------------------
def gelu(x: np.ndarray) -> np.ndarray:
    """tanh-approximated GELU, tf2_gnn/utils/activation.py:7-14."""
    dt = x.dtype.type
    cdf = dt(0.5) * (dt(1.0) + np.tanh(dt(math.sqrt(2 / math.pi)) * (x + dt(0.044715) * x * x * x)))
    return x * cdf


def get_activation_function(name: Optional[str]):
    """tf2_gnn/utils/param_helpers.py:22-42 ("linear" maps to None and therefore raises)."""
    if name is None:
        return None
    name = name.lower()
    table = {
        "linear": None,
        "tanh": np.tanh,
        "relu": lambda x: np.maximum(x, x.dtype.type(0)),
        "leaky_relu": lambda x: np.where(x > 0, x, x.dtype.type(LEAKY_RELU_ALPHA) * x),
        "elu": lambda x: np.where(x > 0, x, np.expm1(np.minimum(x, x.dtype.type(0)))),
        "selu": lambda x: x.dtype.type(SELU_SCALE)
        * np.where(x > 0, x, x.dtype.type(SELU_ALPHA) * np.expm1(np.minimum(x, x.dtype.type(0)))),
        "gelu": gelu,
    }
    fn = table.get(name)
    if fn is None:
        raise ValueError(f"Unknown activation function: {name}")
    return fn


def unsorted_segment_sum(data: np.ndarray, segment_ids: np.ndarray, num_segments: int) -> np.ndarray:
    out = np.zeros((num_segments,) + data.shape[1:], dtype=data.dtype)
    np.add.at(out, segment_ids, data)
    return out


def _segment_counts(segment_ids: np.ndarray, num_segments: int, dtype) -> np.ndarray:
    return np.bincount(segment_ids, minlength=num_segments).astype(dtype)


def unsorted_segment_mean(data, segment_ids, num_segments):
    """tf.math.unsorted_segment_mean: sum / max(count, 1); empty segments give 0 [external TF]."""
    n = np.maximum(_segment_counts(segment_ids, num_segments, data.dtype), 1)
    return unsorted_segment_sum(data, segment_ids, num_segments) / n.reshape((-1,) + (1,) * (data.ndim - 1))


def unsorted_segment_sqrt_n(data, segment_ids, num_segments):
    """tf.math.unsorted_segment_sqrt_n: sum / sqrt(max(count, 1)) [external TF]."""
    n = np.maximum(_segment_counts(segment_ids, num_segments, data.dtype), 1)
    return unsorted_segment_sum(data, segment_ids, num_segments) / np.sqrt(n).reshape(
        (-1,) + (1,) * (data.ndim - 1)
    )


def unsorted_segment_max(data, segment_ids, num_segments):
    """tf.math.unsorted_segment_max: empty segments give the lowest finite value [external TF]."""
    out = np.full((num_segments,) + data.shape[1:], np.finfo(data.dtype).min, dtype=data.dtype)
    np.maximum.at(out, segment_ids, data)
    return out


def get_aggregation_function(name: str):
    """tf2_gnn/utils/param_helpers.py:7-19."""
    table = {
        "sum": unsorted_segment_sum,
        "max": unsorted_segment_max,
        "mean": unsorted_segment_mean,
        "sqrt_n": unsorted_segment_sqrt_n,
    }
    fn = table.get(name)
    if fn is None:
        raise ValueError(f"Unknown aggregation function: {name}")
    return fn


def unsorted_segment_log_softmax(logits, segment_ids, num_segments):
    """dpu_utils.tf2utils.unsorted_segment_log_softmax [external, published algorithm]:
    (x - max_seg[ids]) - log(segment_sum(exp(x - max_seg[ids])))[ids]."""
    max_per_segment = unsorted_segment_max(logits, segment_ids, num_segments)
    recentered = logits - max_per_segment[segment_ids]
    per_segment_sums = unsorted_segment_sum(np.exp(recentered), segment_ids, num_segments)
    with np.errstate(divide="ignore"):
        norm = np.log(per_segment_sums)
    return recentered - norm[segment_ids]


# --------------------------------------------------------------------------------------
# In-degree table (message_passing.py:230-263)
# --------------------------------------------------------------------------------------
def calculate_type_to_num_incoming_edges(num_nodes: int, adjacency_lists: Sequence[np.ndarray],
                                         dtype=np.float32) -> np.ndarray:
    """float [L, V]; c[l, v] = number of type-l edges whose target is v (duplicates counted:
    tf.scatter_nd accumulates, message_passing.py:256-260)."""
    rows = []
    for adj in adjacency_lists:
        adj = np.asarray(adj).reshape(-1, 2)
        targets = adj[:, 1]
        rows.append(unsorted_segment_sum(np.ones(len(targets), dtype=dtype), targets, num_nodes))
    if not rows:
        return np.zeros((0, num_nodes), dtype=dtype)
    return np.stack(rows)


# --------------------------------------------------------------------------------------
# dpu_utils MLP (bias-free Dense stack) and Keras GRUCell [external, published semantics]
# --------------------------------------------------------------------------------------
def mlp_forward(x: np.ndarray, layer_weights: Sequence[np.ndarray]) -> np.ndarray:
    """dpu_utils.tf2utils.MLP(out_size, hidden_layers, use_biases=False, activation_fun=relu):
    hidden Dense layers with ReLU, final Dense linear.  Call sites gnn_edge_mlp.py:76-79,100,
    gnn_film.py:74-78,99-101, rgin.py:81-85,104.  Inference mode (no dropout)."""
    act = x
    for w in layer_weights[:-1]:
        act = np.maximum(act @ w, act.dtype.type(0))
    return act @ layer_weights[-1]


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def gru_cell_forward(x, h, kernel, recurrent_kernel, bias):
    """tf.keras.layers.GRUCell(units=H) with TF2 defaults (reset_after=True, sigmoid/tanh, gate
    order z, r, h; bias [2, 3H]).  Call site ggnn.py:62-66,84-87."""
    H = h.shape[1]
    mx = x @ kernel + bias[0]
    mh = h @ recurrent_kernel + bias[1]
    xz, xr, xh = mx[:, :H], mx[:, H:2 * H], mx[:, 2 * H:]
    rz, rr, rh = mh[:, :H], mh[:, H:2 * H], mh[:, 2 * H:]
    z = _sigmoid(xz + rz).astype(x.dtype)
    r = _sigmoid(xr + rr).astype(x.dtype)
    hh = np.tanh(xh + r * rh)
    return z * h + (1 - z) * hh


# --------------------------------------------------------------------------------------
# The hot path
# --------------------------------------------------------------------------------------
def _gather_per_type(h, adjacency_lists, dtype):
    """message_passing.py:181-206: per type gather h[src], h[tgt], c[l, tgt]."""
    V = h.shape[0]
    c = calculate_type_to_num_incoming_edges(V, adjacency_lists, dtype)
    out = []
    for l, adj in enumerate(adjacency_lists):
        adj = np.asarray(adj).reshape(-1, 2)
        src, tgt = adj[:, 0], adj[:, 1]
        out.append((h[src], h[tgt], c[l][tgt], tgt))
    return out


def _edge_mlp_message(src_states, tgt_states, n_in, mlp_weights, use_target, normalize):
    """gnn_edge_mlp.py:84-107."""
    x = np.concatenate([src_states, tgt_states], axis=1) if use_target else src_states
    m = mlp_forward(x, mlp_weights)
    if normalize:
        dt = m.dtype.type
        m = (dt(1.0) / (n_in + dt(SMALL_NUMBER)))[:, None] * m
    return m


def _default_epilogue(messages_per_type, targets_per_type, V, params):
    """message_passing.py:165-179."""
    H = int(params["hidden_dim"])
    agg = get_aggregation_function(params["aggregation_function"])
    act = get_activation_function(params["message_activation_function"])
    before = bool(params.get("message_activation_before_aggregation", False))
    dtype = messages_per_type[0].dtype if messages_per_type else np.float32
    targets = (np.concatenate(targets_per_type) if targets_per_type else np.zeros((0,), np.int32))
    messages = (np.concatenate(messages_per_type, axis=0) if messages_per_type
                else np.zeros((0, H), dtype))
    if before:
        messages = act(messages)
    out = agg(messages, targets, V)
    if not before:
        out = act(out)
    return out


def message_passing_forward(kind: str, params: Dict[str, Any], weights: Dict[str, Any],
                            node_embeddings: np.ndarray, adjacency_lists: Sequence[np.ndarray],
                            dtype=np.float32) -> np.ndarray:
    """One message-passing layer, inference mode.  `kind` is the lower-cased class name
    (message_passing/__init__.py:10-14).  `weights`:
      edge-MLP family: weights["edge_mlps"][l] = list of layer matrices for type l
      ggnn: + weights["gru_kernel"], ["gru_recurrent_kernel"], ["gru_bias"]
      rgin: + weights["aggr_mlp"] (list of matrices) or None
      gnn_film: + weights["film_mlps"][l] = list of layer matrices ([D,2H] last)
      rgat: weights["edge_kernels"][l] [D,H], weights["edge_attention"][l] [K, 2H/K]
      pass_source_states (test double of test_message_passing.py:11-27): no weights
    """
    kind = kind.lower()
    h = np.asarray(node_embeddings, dtype=dtype)
    adjacency_lists = [np.asarray(a).reshape(-1, 2) for a in adjacency_lists]
    V = h.shape[0]
    H = int(params["hidden_dim"])
    gathered = _gather_per_type(h, adjacency_lists, dtype)
    targets_per_type = [g[3] for g in gathered]

    def cast(ws):
        return [np.asarray(w, dtype=dtype) for w in ws]

    if kind == "pass_source_states":
        msgs = [g[0] for g in gathered]
        return _default_epilogue(msgs, targets_per_type, V, params)

    if kind in ("gnn_edge_mlp", "rgcn", "ggnn", "rgin", "gnn_film"):
        use_target = bool(params["use_target_state_as_input"])
        normalize = bool(params["normalize_by_num_incoming"])
        msgs = []
        for l, (s, t, n, _) in enumerate(gathered):
            m = _edge_mlp_message(s, t, n, cast(weights["edge_mlps"][l]), use_target, normalize)
            if kind == "gnn_film":
                # gnn_film.py:99-107
                film = mlp_forward(t, cast(weights["film_mlps"][l]))
                m = film[:, :H] * m + film[:, H:]
            msgs.append(m)
        if kind in ("gnn_edge_mlp", "rgcn", "gnn_film"):
            return _default_epilogue(msgs, targets_per_type, V, params)
        agg = get_aggregation_function(params["aggregation_function"])
        targets = np.concatenate(targets_per_type) if targets_per_type else np.zeros((0,), np.int32)
        messages = np.concatenate(msgs, axis=0) if msgs else np.zeros((0, H), dtype)
        aggregated = agg(messages, targets, V)
        if kind == "ggnn":
            # ggnn.py:68-89: no activation, GRU(inputs=aggregated, state=h)
            return gru_cell_forward(aggregated, h,
                                    np.asarray(weights["gru_kernel"], dtype=dtype),
                                    np.asarray(weights["gru_recurrent_kernel"], dtype=dtype),
                                    np.asarray(weights["gru_bias"], dtype=dtype)).astype(dtype)
        # rgin.py:88-106
        if weights.get("aggr_mlp") is not None:
            aggregated = mlp_forward(aggregated, cast(weights["aggr_mlp"]))
        return get_activation_function(params["message_activation_function"])(aggregated)

    if kind == "rgat":
        K = int(params["num_heads"])
        d = H // K
        per_head_msgs, scores = [], []
        for l, (s, t, _, _) in enumerate(gathered):
            W = np.asarray(weights["edge_kernels"][l], dtype=dtype)
            a = np.asarray(weights["edge_attention"][l], dtype=dtype)
            ps = (s @ W).reshape(-1, K, d)  # rgat.py:102-105
            pt = (t @ W).reshape(-1, K, d)  # rgat.py:106-109
            cat = np.concatenate([ps, pt], axis=-1)  # rgat.py:111-113
            sc = np.einsum("vki,ki->vk", cat, a)  # rgat.py:115-121
            sc = np.where(sc > 0, sc, dtype(LEAKY_RELU_ALPHA) * sc).astype(dtype)
            per_head_msgs.append(ps)
            scores.append(sc)
        msgs = np.concatenate(per_head_msgs, axis=0) if per_head_msgs else np.zeros((0, K, d), dtype)
        sc = np.concatenate(scores, axis=0) if scores else np.zeros((0, K), dtype)
        targets = np.concatenate(targets_per_type) if targets_per_type else np.zeros((0,), np.int32)
        heads = []
        for k in range(K):  # rgat.py:141-160
            att = np.exp(unsorted_segment_log_softmax(sc[:, k], targets, V)).astype(dtype)
            heads.append(unsorted_segment_sum(att[:, None] * msgs[:, k, :], targets, V))
        out = np.concatenate(heads, axis=-1)
        return get_activation_function(params["message_activation_function"])(out)

    raise ValueError(f"Unknown message passing type: {kind}")


# --------------------------------------------------------------------------------------
# Hyper-parameter defaults (message_passing.py:41-48 and subclasses)
# --------------------------------------------------------------------------------------
def default_hyperparameters(kind: str) -> Dict[str, Any]:
    base = {
        "aggregation_function": "sum",
        "message_activation_function": "relu",
        "message_activation_before_aggregation": False,
        "hidden_dim": 7,
    }
    edge_mlp = dict(base, use_target_state_as_input=True, normalize_by_num_incoming=False,
                    num_edge_MLP_hidden_layers=1)
    kind = kind.lower()
    if kind == "pass_source_states":
        return base
    if kind == "gnn_edge_mlp":
        return edge_mlp
    if kind in ("rgcn", "ggnn"):
        return dict(edge_mlp, use_target_state_as_input=False, normalize_by_num_incoming=True,
                    num_edge_MLP_hidden_layers=0)
    if kind == "gnn_film":
        return dict(edge_mlp, use_target_state_as_input=False, normalize_by_num_incoming=False,
                    num_edge_MLP_hidden_layers=0, film_parameter_MLP_hidden_layers=[])
    if kind == "rgin":
        return dict(edge_mlp, use_target_state_as_input=False, num_edge_MLP_hidden_layers=1,
                    num_aggr_MLP_hidden_layers=None)
    if kind == "rgat":
        return dict(base, num_heads=3)
    raise ValueError(f"Unknown message passing type: {kind}")


def glorot_uniform(rng: np.random.Generator, shape: Tuple[int, ...], dtype=np.float32) -> np.ndarray:
    """Keras default kernel initialiser U(+-sqrt(6/(fan_in+fan_out))) [external Keras]."""
    fan_in, fan_out = shape[-2], shape[-1]
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(dtype)


def make_weights(kind: str, params: Dict[str, Any], D: int, L: int, rng: np.random.Generator,
                 dtype=np.float32) -> Dict[str, Any]:
    """Random weights with the shapes the reference builds (gnn_edge_mlp.py:64-82,
    ggnn.py:62-66, rgat.py:68-89, gnn_film.py:67-81, rgin.py:77-86)."""
    kind = kind.lower()
    H = int(params["hidden_dim"])
    w: Dict[str, Any] = {}
    if kind == "pass_source_states":
        return w
    if kind == "rgat":
        K = int(params["num_heads"])
        w["edge_kernels"] = [glorot_uniform(rng, (D, H), dtype) for _ in range(L)]
        w["edge_attention"] = [glorot_uniform(rng, (K, 2 * (H // K)), dtype) for _ in range(L)]
        return w
    in_dim = 2 * D if params["use_target_state_as_input"] else D
    n_hidden = int(params["num_edge_MLP_hidden_layers"])
    sizes = [in_dim] + [H] * n_hidden + [H]
    w["edge_mlps"] = [[glorot_uniform(rng, (sizes[i], sizes[i + 1]), dtype)
                       for i in range(len(sizes) - 1)] for _ in range(L)]
    if kind == "ggnn":
        w["gru_kernel"] = glorot_uniform(rng, (D, 3 * H), dtype)
        w["gru_recurrent_kernel"] = glorot_uniform(rng, (H, 3 * H), dtype)
        w["gru_bias"] = rng.uniform(-0.1, 0.1, size=(2, 3 * H)).astype(dtype)
    if kind == "rgin":
        n_aggr = params.get("num_aggr_MLP_hidden_layers")
        if n_aggr is not None:
            s = [H] + [H] * int(n_aggr) + [H]
            w["aggr_mlp"] = [glorot_uniform(rng, (s[i], s[i + 1]), dtype) for i in range(len(s) - 1)]
        else:
            w["aggr_mlp"] = None
    if kind == "gnn_film":
        hidden = list(params.get("film_parameter_MLP_hidden_layers", []))
        s = [D] + hidden + [2 * H]
        w["film_mlps"] = [[glorot_uniform(rng, (s[i], s[i + 1]), dtype)
                           for i in range(len(s) - 1)] for _ in range(L)]
    return w


# --------------------------------------------------------------------------------------
# Graph readout and global exchange (nodes_to_graph_representation.py:170-229,
# graph_global_exchange.py:83-183).  PARITY UNPINNED (no reference test holds a value; dpu_utils.MLP and
# unsorted_segment_softmax are external) until tests/golden/tf_layers_golden.json exists (tools/gen_tf_golden.py).
# --------------------------------------------------------------------------------------
def dense_mlp_forward(x: np.ndarray, kernels: Sequence[np.ndarray], biases: Optional[Sequence[np.ndarray]] = None,
                      activation=None) -> np.ndarray:
    """dpu_utils.tf2utils.MLP with an arbitrary hidden activation and optional biases (the readout MLPs:
    nodes_to_graph_representation.py:128-148): hidden Dense layers with `activation`, linear output layer."""
    act = activation or (lambda v: np.maximum(v, v.dtype.type(0)))
    cur = x
    n = len(kernels)
    for i, w in enumerate(kernels):
        cur = cur @ w
        if biases is not None and biases[i] is not None:
            cur = cur + biases[i]
        if i < n - 1:
            cur = act(cur)
    return cur


def unsorted_segment_softmax(logits, segment_ids, num_segments):
    """dpu_utils.tf2utils.unsorted_segment_softmax = exp(unsorted_segment_log_softmax) [external]."""
    return np.exp(unsorted_segment_log_softmax(logits, segment_ids, num_segments))


def weighted_sum_graph_representation(node_embeddings, node_to_graph_map, num_graphs, weights: Dict[str, Any],
                                      graph_representation_size: int, num_heads: int, weighting_fun: str = "softmax",
                                      scoring_activation: str = "relu", transformation_activation: str = "relu",
                                      lower_bound: Optional[float] = None, upper_bound: Optional[float] = None,
                                      dtype=np.float32) -> np.ndarray:
    """WeightedSumGraphRepresentation.call, inference mode (nodes_to_graph_representation.py:170-229).
    weights: {"scoring_mlp": [kernels], "transformation_mlp": [kernels], optional "scoring_biases",
    "transformation_biases"}.  Returns [num_graphs, GD] (the reference's tf.math.segment_sum returns
    max(id)+1 rows: identical whenever the last graph of the batch has a node, which graph_dataset.py guarantees)."""
    x = np.asarray(node_embeddings, dtype=dtype)
    ids = np.asarray(node_to_graph_map).astype(np.int64)
    G, GD, K = int(num_graphs), int(graph_representation_size), int(num_heads)
    weighting_fun = weighting_fun.lower()

    def cast(ws):
        return None if ws is None else [None if w is None else np.asarray(w, dtype=dtype) for w in ws]

    w = None
    if weighting_fun not in ("none", "average"):                                   # :172-188
        scores = dense_mlp_forward(x, cast(weights["scoring_mlp"]), cast(weights.get("scoring_biases")),
                                   get_activation_function(scoring_activation))   # [V, K]
        if weighting_fun == "sigmoid":
            w = _sigmoid(scores).astype(dtype)
        elif weighting_fun == "softmax":
            w = np.stack([unsorted_segment_softmax(scores[:, k], ids, G) for k in range(K)], axis=1).astype(dtype)
        else:
            raise ValueError()
    t_act = get_activation_function(transformation_activation)
    reprs = t_act(dense_mlp_forward(x, cast(weights["transformation_mlp"]),
                                    cast(weights.get("transformation_biases")), t_act))  # :191-193
    if lower_bound is not None:
        reprs = np.maximum(reprs, dtype(lower_bound))
    if upper_bound is not None:
        reprs = np.minimum(reprs, dtype(upper_bound))
    if weighting_fun == "none":                                                     # :204-210
        return unsorted_segment_sum(reprs, ids, G)
    if weighting_fun == "average":                                                  # :211-217
        cnt = np.maximum(np.bincount(ids, minlength=G), 1).astype(dtype)
        return unsorted_segment_sum(reprs, ids, G) / cnt[:, None]
    reprs = reprs.reshape(-1, K, GD // K) * w[:, :, None]                           # :219-220
    return unsorted_segment_sum(reprs.reshape(-1, GD), ids, G)                      # :222-227


def graph_global_exchange(mode: str, node_embeddings, node_to_graph_map, num_graphs, weights: Dict[str, Any],
                          hidden_dim: int, num_heads: int, weighting_fun: str = "softmax", dtype=np.float32):
    """GraphGlobal{Mean,GRU,MLP}Exchange.call, inference mode (graph_global_exchange.py:83-183).
    weights: the readout's {"scoring_mlp", "transformation_mlp"} + gru: "gru_kernel", "gru_recurrent_kernel",
    "gru_bias"; mlp: "mlp" (kernels of MLP(out_size=H) on [graph repr || node state])."""
    x = np.asarray(node_embeddings, dtype=dtype)
    ids = np.asarray(node_to_graph_map).astype(np.int64)
    g = weighted_sum_graph_representation(x, ids, num_graphs, weights, hidden_dim, num_heads, weighting_fun,
                                          dtype=dtype)                              # :84-92
    per_node = g[ids]                                                               # :94-96 gather_dense_gradient
    mode = mode.lower()
    if mode == "mean":
        return (x + per_node) / dtype(2)                                            # :124
    if mode == "gru":                                                               # :147-152
        return gru_cell_forward(per_node, x, np.asarray(weights["gru_kernel"], dtype=dtype),
                                np.asarray(weights["gru_recurrent_kernel"], dtype=dtype),
                                np.asarray(weights["gru_bias"], dtype=dtype)).astype(dtype)
    if mode == "mlp":                                                               # :176-181
        return mlp_forward(np.concatenate([per_node, x], axis=-1), [np.asarray(w, dtype=dtype) for w in weights["mlp"]])
    raise ValueError(f"Unknown global_exchange_mode mode {mode}")


def make_exchange_weights(mode: str, hidden_dim: int, num_heads: int, rng: np.random.Generator,
                          weighting_fun: str = "softmax", dtype=np.float32) -> Dict[str, Any]:
    """Shapes as GraphGlobalExchange.build creates them (graph_global_exchange.py:46-58,138-140,167-169):
    scoring MLP [H -> H -> num_heads], transformation MLP [H -> 128 -> H] (the class default layer list), no biases."""
    H = hidden_dim
    w: Dict[str, Any] = {"transformation_mlp": [glorot_uniform(rng, (H, 128), dtype), glorot_uniform(rng, (128, H), dtype)]}
    if weighting_fun.lower() in ("softmax", "sigmoid"):
        w["scoring_mlp"] = [glorot_uniform(rng, (H, H), dtype), glorot_uniform(rng, (H, num_heads), dtype)]
    if mode.lower() == "gru":
        w["gru_kernel"] = glorot_uniform(rng, (H, 3 * H), dtype)
        w["gru_recurrent_kernel"] = glorot_uniform(rng, (H, 3 * H), dtype)
        w["gru_bias"] = rng.uniform(-0.1, 0.1, size=(2, 3 * H)).astype(dtype)
    if mode.lower() == "mlp":
        w["mlp"] = [glorot_uniform(rng, (2 * H, H), dtype), glorot_uniform(rng, (H, H), dtype)]
    return w


# --------------------------------------------------------------------------------------
# GNN stack, inference mode (gnn.py:276-329), including the global exchange layers.
# --------------------------------------------------------------------------------------
def layer_norm(x: np.ndarray, gamma: np.ndarray, beta: np.ndarray, epsilon: float = 1e-3) -> np.ndarray:
    """tf.keras.layers.LayerNormalization() defaults (axis=-1, epsilon=1e-3) [external Keras]."""
    mean = x.mean(axis=-1, keepdims=True)
    var = ((x - mean) ** 2).mean(axis=-1, keepdims=True)
    return (x - mean) / np.sqrt(var + x.dtype.type(epsilon)) * gamma + beta


def gnn_forward(params: Dict[str, Any], weights: Dict[str, Any], node_features: np.ndarray,
                adjacency_lists: Sequence[np.ndarray], dtype=np.float32,
                node_to_graph_map: Optional[np.ndarray] = None, num_graphs: Optional[int] = None):
    """weights: {"initial_projection": [F,H], "mp": [per-layer message-passing weight dicts],
    "dense": {layer_idx: [H,H]}, "layernorm": [(gamma, beta) per layer], "exchange": {layer_idx: exchange weights}}.
    Returns (final representations, tuple of all representations) like GNN._internal_call."""
    kind = params["message_calculation_class"].lower()
    act_init = get_activation_function(params["initial_node_representation_activation"])
    act_dense = get_activation_function(params["dense_intermediate_layer_activation"])
    x = np.asarray(node_features, dtype=dtype) @ np.asarray(weights["initial_projection"], dtype=dtype)
    cur = act_init(x) if act_init is not None else x                      # gnn.py:279
    last = cur
    all_reps = [cur]
    for i in range(int(params["num_layers"])):
        if i % int(params["residual_every_num_layers"]) == 0:             # gnn.py:291-296
            tmp = cur
            if i > 0:
                cur = (cur + last) / dtype(2)
            last = tmp
        cur = message_passing_forward(kind, params, weights["mp"][i], cur, adjacency_lists, dtype=dtype)
        all_reps.append(cur)                                              # gnn.py:305
        if i and i % int(params["global_exchange_every_num_layers"]) == 0:   # gnn.py:307-315
            ex = weights["exchange"][i] if i in weights["exchange"] else weights["exchange"][str(i)]
            cur = graph_global_exchange(params["global_exchange_mode"], cur, node_to_graph_map, int(num_graphs), ex,
                                        int(params["hidden_dim"]), int(params["global_exchange_num_heads"]),
                                        params["global_exchange_weighting_fun"], dtype=dtype).astype(dtype)
        if params["use_inter_layer_layernorm"]:                           # gnn.py:317-321
            g, b = weights["layernorm"][i]
            cur = layer_norm(cur, np.asarray(g, dtype=dtype), np.asarray(b, dtype=dtype)).astype(dtype)
        if i % int(params["dense_every_num_layers"]) == 0:                # gnn.py:324-327
            d = weights["dense"][i] if i in weights["dense"] else weights["dense"][str(i)]
            y = cur @ np.asarray(d, dtype=dtype)
            cur = act_dense(y) if act_dense is not None else y
    return cur, tuple(all_reps)
