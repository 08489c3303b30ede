"""Generate tests/golden/*.json.  Run in the BUILD container only (needs /root/reference).

  python oracle/gen_golden.py

1. message_passing_golden.json — the golden vectors the reference's own tests hold for the
   hot path, transcribed by hand (TensorFlow is absent, so the tests cannot be executed):
     * tf2_gnn/test/layers/test_message_passing.py:35-71  (4 gather/sum/relu cases)
     * tf2_gnn/layers/message_passing/message_passing.py:238-249 (in-degree doctest)
2. process_adjacency_lists_golden.json — produced by EXECUTING the reference's
   tf2_gnn/data/utils.py (pure numpy; loaded by file path so tf2_gnn/__init__ and its
   TensorFlow imports are never touched) on the 8 inputs of
   tf2_gnn/test/data/test_utils.py:50-115 plus seeded random inputs; the expected outputs
   written in that test file are stored next to the executed outputs and must agree.
3. batch_assembly_golden.json — produced by EXECUTING the reference's own minibatch assembly,
   tf2_gnn/data/graph_dataset.py:161-246 (graph_batch_iterator_from_graph_iterator, _batch_would_be_too_full,
   _add_graph_to_batch, _finalise_batch).  The module imports tensorflow and dpu_utils only for its tf.data
   wrapper and type annotations; both are replaced by attribute-swallowing stubs for the import, the batching
   code itself is pure numpy and runs unmodified on seeded random graphs.
"""
import importlib.util
import json
import os

import numpy as np

REF = "/root/reference/tf2_gnn"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _load_reference_data_utils():
    spec = importlib.util.spec_from_file_location("_ref_data_utils", os.path.join(REF, "data", "utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def message_passing_golden():
    cases = [
        dict(name="node0_to_node1", source="test_message_passing.py:36-43",
             node_embeddings=[[1, 2, 3], [2, 4, 5]], adjacency_lists=[[[0, 1]]],
             aggregated_states=[[0, 0, 0], [1, 2, 3]]),
        dict(name="node1_to_node0_relu_zeroes_negative", source="test_message_passing.py:44-51",
             node_embeddings=[[1, 2, 3], [2, -4, 5]], adjacency_lists=[[[1, 0]]],
             aggregated_states=[[2, 0, 5], [0, 0, 0]]),
        dict(name="two_messages_summed_in_node2", source="test_message_passing.py:52-59",
             node_embeddings=[[1, 2, 3], [2, 4, 5], [0, -7, -4]],
             adjacency_lists=[[[1, 0], [0, 1], [0, 2], [1, 2]]],
             aggregated_states=[[2, 4, 5], [1, 2, 3], [3, 6, 8]]),
        dict(name="self_loop_in_second_edge_type", source="test_message_passing.py:60-70",
             node_embeddings=[[1, 2, 3], [2, 4, 5], [0, -7, -4]],
             adjacency_lists=[[[1, 0], [0, 1], [0, 2], [1, 2]], [[2, 2]]],
             aggregated_states=[[2, 4, 5], [1, 2, 3], [3, 0, 4]]),
    ]
    indegree = dict(source="message_passing.py:238-249", num_nodes=5,
                    adjacency_lists=[[[0, 1], [2, 4], [2, 4]], [[2, 3], [2, 4]], [[3, 1]]],
                    type_to_num_incoming_edges=[[0, 1, 0, 0, 2], [0, 0, 0, 1, 1], [0, 1, 0, 0, 0]])
    return dict(pass_source_states=cases, in_degree_doctest=indegree,
                params=dict(message_activation_function="relu", aggregation_function="sum"))


def process_adjacency_lists_golden():
    ref = _load_reference_data_utils()

    def tinput(add_self, tie, two=False, slt=0):
        return dict(adjacency_lists=[[[0, 1]], [[1, 2]]] if two else [[[0, 1], [1, 2]]],
                    num_nodes=3, add_self_loop_edges=add_self, tie_fwd_bkwd_edges=tie,
                    self_loop_edge_type=slt)

    # (input, expected adjacency lists, expected in-degree) from test_utils.py:50-115
    transcribed = [
        (tinput(False, False), [[[0, 1], [1, 2]], [[1, 0], [2, 1]]], [[0, 1, 1], [1, 1, 0]]),
        (tinput(False, True), [[[0, 1], [1, 2], [1, 0], [2, 1]]], [[1, 2, 1]]),
        (tinput(True, False), [[[0, 0], [1, 1], [2, 2]], [[0, 1], [1, 2]], [[1, 0], [2, 1]]],
         [[1, 1, 1], [0, 1, 1], [1, 1, 0]]),
        (tinput(True, True), [[[0, 0], [1, 1], [2, 2]], [[0, 1], [1, 2], [1, 0], [2, 1]]],
         [[1, 1, 1], [1, 2, 1]]),
        (tinput(True, False, slt=-1), [[[0, 1], [1, 2]], [[1, 0], [2, 1]], [[0, 0], [1, 1], [2, 2]]],
         [[0, 1, 1], [1, 1, 0], [1, 1, 1]]),
        (tinput(True, True, slt=-1), [[[0, 1], [1, 2], [1, 0], [2, 1]], [[0, 0], [1, 1], [2, 2]]],
         [[1, 2, 1], [1, 1, 1]]),
        (tinput(False, [0], two=True), [[[0, 1], [1, 0]], [[1, 2]], [[2, 1]]],
         [[1, 1, 0], [0, 0, 1], [0, 1, 0]]),
        (tinput(False, [1], two=True), [[[0, 1]], [[1, 2], [2, 1]], [[1, 0]]],
         [[0, 1, 0], [0, 1, 1], [1, 0, 0]]),
    ]
    cases = []

    def run(inp):
        tied = ref.get_tied_edge_types(inp["tie_fwd_bkwd_edges"], len(inp["adjacency_lists"]))
        adj_in = [[tuple(e) for e in a] for a in inp["adjacency_lists"]]
        adj, cnt = ref.process_adjacency_lists(
            adjacency_lists=adj_in, num_nodes=inp["num_nodes"],
            add_self_loop_edges=inp["add_self_loop_edges"], tied_fwd_bkwd_edge_types=tied,
            self_loop_edge_type=inp["self_loop_edge_type"])
        return [a.tolist() for a in adj], cnt.tolist()

    for inp, exp_adj, exp_cnt in transcribed:
        got_adj, got_cnt = run(inp)
        assert got_adj == exp_adj and got_cnt == [[float(x) for x in r] for r in exp_cnt], inp
        cases.append(dict(source="test_utils.py:50-115 (transcribed) == executed reference",
                          input=inp, adjacency_lists=got_adj, type_to_num_incoming_edges=got_cnt))
    rng = np.random.default_rng(0)
    for i in range(12):
        n = int(rng.integers(1, 30))
        n_types = int(rng.integers(1, 5))
        adj = [rng.integers(0, n, size=(int(rng.integers(0, 40)), 2)).tolist() for _ in range(n_types)]
        tie_choice = [True, False, sorted(set(rng.integers(0, n_types, size=2).tolist()))][i % 3]
        add_self = bool(i % 2)
        inp = dict(adjacency_lists=adj, num_nodes=n, add_self_loop_edges=add_self,
                   tie_fwd_bkwd_edges=tie_choice,
                   self_loop_edge_type=int(rng.integers(-2, 2)) if add_self else 0)
        got_adj, got_cnt = run(inp)
        cases.append(dict(source="executed reference data/utils.py, seeded random input",
                          input=inp, adjacency_lists=got_adj, type_to_num_incoming_edges=got_cnt))
    return dict(cases=cases)


class _Stub:
    """Stands in for `tensorflow` / `dpu_utils.utils` while the reference module is imported (annotations only)."""

    def __getattr__(self, name):
        return self

    def __call__(self, *a, **k):
        return self

    def __getitem__(self, k):
        return self


def _load_reference_graph_dataset():
    import sys
    import types
    saved = {k: sys.modules.get(k) for k in ("tensorflow", "dpu_utils", "dpu_utils.utils")}
    try:
        tf_stub = _Stub()
        du = types.ModuleType("dpu_utils")
        duu = types.ModuleType("dpu_utils.utils")
        duu.RichPath = _Stub()
        duu.DoubleBufferedIterator = _Stub()
        du.utils = duu
        sys.modules["tensorflow"] = tf_stub
        sys.modules["dpu_utils"] = du
        sys.modules["dpu_utils.utils"] = duu
        spec = importlib.util.spec_from_file_location("_ref_graph_dataset", os.path.join(REF, "data", "graph_dataset.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def batch_assembly_golden():
    ref = _load_reference_graph_dataset()

    class _Dataset(ref.GraphDataset):   # the abstract hooks are not touched by the batching code
        def __init__(self, params, n_types, feat_dim):
            super().__init__(params)
            self._n_types, self._feat_dim = n_types, feat_dim

        num_edge_types = property(lambda self: self._n_types)
        node_feature_shape = property(lambda self: (self._feat_dim,))

        def load_data(self, path, folds_to_load=None):
            raise NotImplementedError

        def load_data_from_list(self, datapoints, target_fold=None):
            raise NotImplementedError

        def _graph_iterator(self, data_fold):
            raise NotImplementedError

    cases = []
    for seed, (G, T, max_n, F, max_nodes_per_batch) in enumerate([(1, 1, 4, 2, 100), (7, 2, 6, 3, 12), (25, 3, 9, 4, 40),
                                                                 (12, 4, 29, 15, 100)]):
        rng = np.random.default_rng(100 + seed)
        graphs = []
        for g in range(G):
            n = int(rng.integers(1, max_n + 1))
            adj = []
            for t in range(T):
                e = 0 if (t == 1 and g % 3 == 0) else int(rng.integers(0, 2 * n + 1))
                adj.append(rng.integers(0, n, size=(e, 2)).astype(np.int32))
            graphs.append(dict(node_features=rng.integers(-3, 4, size=(n, F)).astype(np.float32), adjacency_lists=adj))
        ds = _Dataset({"max_nodes_per_batch": max_nodes_per_batch}, T, F)
        samples = [ref.GraphSample(adjacency_lists=g["adjacency_lists"], type_to_node_to_num_inedges=None,
                                   node_features=g["node_features"]) for g in graphs]
        batches = []
        first = 0
        for feats, _labels in ds.graph_batch_iterator_from_graph_iterator(iter(samples)):
            nb = int(feats["num_graphs_in_batch"])
            batches.append(dict(
                graph_ids=list(range(first, first + nb)),
                node_features=np.asarray(feats["node_features"], np.float32).tolist(),
                node_to_graph_map=np.asarray(feats["node_to_graph_map"]).astype(int).tolist(),
                num_graphs_in_batch=nb,
                adjacency_lists=[np.asarray(feats[f"adjacency_list_{t}"]).astype(int).reshape(-1, 2).tolist()
                                 for t in range(T)],
                adjacency_dtypes=[str(np.asarray(feats[f"adjacency_list_{t}"]).dtype) for t in range(T)]))
            first += nb
        cases.append(dict(
            source="executed tf2_gnn/data/graph_dataset.py:161-246", num_edge_types=T, feature_dim=F,
            max_nodes_per_batch=max_nodes_per_batch,
            graphs=[dict(node_features=g["node_features"].tolist(),
                         adjacency_lists=[a.astype(int).tolist() for a in g["adjacency_lists"]]) for g in graphs],
            batches=batches))
    return dict(cases=cases)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "message_passing_golden.json"), "w") as f:
        json.dump(message_passing_golden(), f, indent=1)
    with open(os.path.join(OUT, "process_adjacency_lists_golden.json"), "w") as f:
        json.dump(process_adjacency_lists_golden(), f)
    print("wrote golden fixtures to", os.path.normpath(OUT))
    with open(os.path.join(OUT, "batch_assembly_golden.json"), "w") as f:
        json.dump(batch_assembly_golden(), f)
