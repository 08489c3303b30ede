#!/usr/bin/env python
"""Pin the [external] arithmetic of the oracle against the REAL reference (run this where TensorFlow exists).

    pip install tensorflow dpu-utils
    python tools/gen_tf_golden.py --reference /path/to/tf2-gnn [--out tests/golden/tf_layers_golden.json]

This build image has neither TensorFlow nor dpu_utils, so the semantics the oracle restates "from
the published algorithm" — Keras GRUCell (gate order z,r,h; reset_after=True; bias [2,3H]), dpu_utils.tf2utils.MLP
(hidden width for an integer `hidden_layers`, bias-free Dense stack, ReLU), unsorted_segment_log_softmax /
unsorted_segment_softmax, tf.nn.leaky_relu alpha, tf.math.unsorted_segment_{mean,max,sqrt_n}, LayerNormalization
epsilon — cannot be falsified here.  This script closes that gap on any machine that has them: it runs the
UNMODIFIED reference layers (tf2_gnn.layers.message_passing.{RGCN,GGNN,RGAT,RGIN,GNN_Edge_MLP,GNN_FiLM}, the GNN stack
with global exchange, WeightedSumGraphRepresentation) on seeded inputs, reads the weights back out of the layer
objects, and writes inputs + weights + outputs to one JSON file.

tests/test_tf_golden.py consumes the file when it is present:
  * CPU (`-m "not gpu"`): the numpy oracle must reproduce every recorded output within 1e-5 (norm-wise) — this is
    what turns "parity unpinned" into "pinned" for the rows of SURVEY.md §8c marked [external];
  * GPU (`-m gpu`): the CUDA path must reproduce them too.

Weights are exported by walking the layer objects (never by variable name), in the oracle's dict layout
(oracle/message_passing_oracle.py: message_passing_forward docstring).
"""
import argparse
import json
import os
import sys

import numpy as np


def _np(x):
    return np.asarray(x).tolist()


def mlp_kernels(mlp):
    """Kernels of a dpu_utils.tf2utils.MLP in application order (hidden layers first, output layer last).  The MLP keeps
    its Dense layers in a list; trainable_variables follows creation order = application order."""
    ks = [v.numpy() for v in mlp.trainable_variables]
    for a, b in zip(ks[:-1], ks[1:]):
        assert a.ndim == 2 and b.ndim == 2 and a.shape[1] == b.shape[0], "MLP has biases or an unexpected layout"
    return ks


def export_weights(kind, layer):
    w = {}
    if kind == "rgat":
        w["edge_kernels"] = [_np(d.kernel.numpy()) for d in layer._edge_type_to_message_computation_layer]
        w["edge_attention"] = [_np(a.numpy()) for a in layer._edge_type_to_attention_parameters]
        return w
    w["edge_mlps"] = [[_np(k) for k in mlp_kernels(m)] for m in layer._edge_type_mlps]
    if kind == "ggnn":
        kernel, recurrent, bias = layer._recurrent_unit.get_weights()
        w["gru_kernel"], w["gru_recurrent_kernel"], w["gru_bias"] = _np(kernel), _np(recurrent), _np(bias)
        w["gru_bias_shape"] = list(np.asarray(bias).shape)      # [2, 3H] iff reset_after=True
    if kind == "rgin":
        w["aggr_mlp"] = ([_np(k) for k in mlp_kernels(layer._aggregation_mlp)]
                         if layer._aggregation_mlp is not None else None)
    if kind == "gnn_film":
        w["film_mlps"] = [[_np(k) for k in mlp_kernels(m)] for m in layer._edge_type_film_layer_computations]
    return w


def random_graph(rng, V, L, E):
    adjs = []
    for l in range(L):
        a = rng.integers(0, V, size=(E, 2)).astype(np.int32)
        if l == 0:
            a[: E // 3, 1] = V // 2          # a small hub: exercises the segment softmax / max over many edges
        adjs.append(a)
    adjs.append(np.zeros((0, 2), np.int32))  # an empty edge type (graph_dataset.py:244)
    return adjs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("TF2_GNN_REFERENCE", "/root/reference"))
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                  "tests", "golden", "tf_layers_golden.json"))
    args = ap.parse_args()
    sys.path.insert(0, args.reference)
    import tensorflow as tf
    import dpu_utils
    from tf2_gnn.layers import GNN, GNNInput, WeightedSumGraphRepresentation, NodesToGraphRepresentationInput
    from tf2_gnn.layers.message_passing import MessagePassingInput, get_message_passing_class

    tf.random.set_seed(0)
    rng = np.random.default_rng(0)
    V, D, L, E = 60, 16, 3, 150
    doc = {"_generator": "tools/gen_tf_golden.py", "tensorflow": tf.__version__,
           "dpu_utils": getattr(dpu_utils, "__version__", "unknown"), "layers": [], "gnn": [], "readout": []}

    layer_cases = [
        ("rgcn", {}), ("rgcn", {"aggregation_function": "mean", "message_activation_function": "tanh"}),
        ("rgcn", {"aggregation_function": "sqrt_n", "message_activation_function": "leaky_relu"}),
        ("rgcn", {"aggregation_function": "max", "message_activation_function": "elu"}),
        ("rgcn", {"message_activation_function": "gelu", "message_activation_before_aggregation": True}),
        ("gnn_edge_mlp", {}), ("gnn_edge_mlp", {"num_edge_MLP_hidden_layers": 2, "normalize_by_num_incoming": True}),
        ("ggnn", {"hidden_dim": D}), ("ggnn", {"hidden_dim": D, "normalize_by_num_incoming": False}),
        ("rgin", {}), ("rgin", {"num_aggr_MLP_hidden_layers": 1}),
        ("gnn_film", {}), ("gnn_film", {"use_target_state_as_input": True, "normalize_by_num_incoming": True}),
        ("rgat", {"hidden_dim": 12, "num_heads": 3}), ("rgat", {"hidden_dim": 16, "num_heads": 4,
                                                                "message_activation_function": "selu"}),
    ]
    for kind, extra in layer_cases:
        cls = get_message_passing_class(kind)
        params = cls.get_default_hyperparameters()
        params["hidden_dim"] = 12
        params.update(extra)
        adjs = random_graph(rng, V, L, E)
        h = rng.uniform(-1, 1, (V, D)).astype(np.float32)
        layer = cls(params)
        out = layer(MessagePassingInput(tf.constant(h), tuple(tf.constant(a) for a in adjs)), training=False)
        if kind == "ggnn":
            # a non-zero GRU bias, so that the [2,3H] split between input and recurrent bias is pinned too
            k, r, b = layer._recurrent_unit.get_weights()
            layer._recurrent_unit.set_weights([k, r, rng.uniform(-0.2, 0.2, b.shape).astype(np.float32)])
            out = layer(MessagePassingInput(tf.constant(h), tuple(tf.constant(a) for a in adjs)), training=False)
        doc["layers"].append({"kind": kind, "params": params, "node_embeddings": _np(h),
                              "adjacency_lists": [_np(a) for a in adjs], "weights": export_weights(kind, layer),
                              "output": _np(out.numpy())})

    # ---- GNN stack incl. global exchange (gnn.py:276-329, graph_global_exchange.py) and LayerNorm ----
    for mode in ("gru", "mlp", "mean"):
        for weighting in ("softmax", "sigmoid"):
            params = GNN.get_default_hyperparameters("rgcn")
            params.update(hidden_dim=16, num_layers=4, global_exchange_mode=mode, global_exchange_every_num_layers=2,
                          global_exchange_weighting_fun=weighting, global_exchange_num_heads=4,
                          use_inter_layer_layernorm=True)
            adjs = random_graph(rng, V, 2, E)
            feats = rng.uniform(-1, 1, (V, 10)).astype(np.float32)
            n2g = np.sort(rng.integers(0, 5, size=V)).astype(np.int32)
            gnn = GNN(params)
            inp = GNNInput(tf.constant(feats), tuple(tf.constant(a) for a in adjs), tf.constant(n2g), tf.constant(5))
            out, all_reps = gnn(inp, training=False, return_all_representations=True)
            weights = {"initial_projection": _np(gnn._initial_projection_layer.kernel.numpy()),
                       "mp": [export_weights("rgcn", l) for l in gnn._mp_layers],
                       "dense": {k: _np(d.kernel.numpy()) for k, d in gnn._dense_layers.items()},
                       "layernorm": [[_np(ln.gamma.numpy()), _np(ln.beta.numpy())] for ln in gnn._inter_layer_layernorms],
                       "exchange": {}}
            for k, ex in gnn._global_exchange_layers.items():
                rep = ex._node_to_graph_representation_layer
                e = {"scoring_mlp": [_np(x) for x in mlp_kernels(rep._scoring_mlp)],
                     "transformation_mlp": [_np(x) for x in mlp_kernels(rep._transformation_mlp)]}
                if mode == "gru":
                    kk, rr, bb = ex._gru_cell.get_weights()
                    e.update(gru_kernel=_np(kk), gru_recurrent_kernel=_np(rr), gru_bias=_np(bb))
                if mode == "mlp":
                    e["mlp"] = [_np(x) for x in mlp_kernels(ex._mlp)]
                weights["exchange"][k] = e
            doc["gnn"].append({"params": params, "node_features": _np(feats), "adjacency_lists": [_np(a) for a in adjs],
                               "node_to_graph_map": _np(n2g), "num_graphs": 5, "weights": weights,
                               "output": _np(out.numpy()), "all_representations": [_np(r.numpy()) for r in all_reps]})

    # ---- graph readout (nodes_to_graph_representation.py:170-229) ----
    for weighting in ("softmax", "sigmoid", "average", "none"):
        x = rng.uniform(-1, 1, (V, 16)).astype(np.float32)
        n2g = np.sort(rng.integers(0, 6, size=V)).astype(np.int32)
        layer = WeightedSumGraphRepresentation(graph_representation_size=12, num_heads=3, weighting_fun=weighting,
                                               scoring_mlp_layers=[16], transformation_mlp_layers=[20])
        out = layer(NodesToGraphRepresentationInput(tf.constant(x), tf.constant(n2g), tf.constant(6)), training=False)
        w = {"transformation_mlp": [_np(k) for k in mlp_kernels(layer._transformation_mlp)]}
        if weighting in ("softmax", "sigmoid"):
            w["scoring_mlp"] = [_np(k) for k in mlp_kernels(layer._scoring_mlp)]
        doc["readout"].append({"weighting_fun": weighting, "num_heads": 3, "graph_representation_size": 12,
                               "node_embeddings": _np(x), "node_to_graph_map": _np(n2g), "num_graphs": 6, "weights": w,
                               "output": _np(out.numpy())})

    with open(args.out, "w") as f:
        json.dump(doc, f)
    print(f"wrote {args.out}: {len(doc['layers'])} layer cases, {len(doc['gnn'])} GNN stacks, "
          f"{len(doc['readout'])} readouts (tensorflow {tf.__version__})")


if __name__ == "__main__":
    main()
