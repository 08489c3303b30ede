"""Host-side logic of the on-device batch builder (no GPU): processed-layout sizes against the pinned adjacency oracle
and the reference's golden cases, the tie / type-count helpers, the greedy batching rule."""
import json
import os
from ctypes import byref, c_int32, c_int64

import numpy as np
import pytest

from oracle import adjacency_oracle as ao
from tf2_gnn_b200 import _ffi
from tf2_gnn_b200.data.graph_store import greedy_batches
from tf2_gnn_b200.data.utils import compute_number_of_edge_types, get_tied_edge_types

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "process_adjacency_lists_golden.json")


def _sizes(E_fwd, V, add_self, tied_set, self_type):
    T = len(E_fwd)
    E = (c_int64 * max(T, 1))(*E_fwd)
    tied = (c_int32 * max(T, 1))(*[1 if t in tied_set else 0 for t in range(T)])
    L = c_int32(0)
    out = (c_int64 * (2 * T + 1))()
    rc = _ffi.lib().tfgnn_b200_process_adjacency_sizes(E, T, V, int(add_self), tied, self_type, out, byref(L))
    return rc, [int(out[i]) for i in range(L.value)]


def test_sizes_match_reference_golden_cases():
    with open(GOLDEN) as f:
        cases = json.load(f)["cases"]
    assert len(cases) >= 8
    for c in cases:
        inp = c["input"]
        T = len(inp["adjacency_lists"])
        tied = ao.get_tied_edge_types(inp["tie_fwd_bkwd_edges"], T)
        E_fwd = [len(np.asarray(a).reshape(-1, 2)) for a in inp["adjacency_lists"]]
        rc, sizes = _sizes(E_fwd, inp["num_nodes"], inp["add_self_loop_edges"], tied, inp["self_loop_edge_type"])
        assert rc == 0
        assert sizes == [len(np.asarray(a).reshape(-1, 2)) for a in c["adjacency_lists"]]
        assert len(sizes) == ao.compute_number_of_edge_types(tied, T, inp["add_self_loop_edges"])


@pytest.mark.parametrize("self_type", [0, 1, 3, -1, -2, -4])
def test_sizes_match_oracle_for_self_loop_slots(self_type):
    rng = np.random.default_rng(3)
    adjs = [rng.integers(0, 7, size=(n, 2)).astype(np.int32) for n in (5, 0, 9)]
    tied = {1}
    expect, _ = ao.process_adjacency_lists(adjs, 7, True, tied, self_type)
    rc, sizes = _sizes([5, 0, 9], 7, True, tied, self_type)
    assert rc == 0 and sizes == [len(a) for a in expect]


def test_self_loop_slot_out_of_range_is_rejected_like_the_reference_assert():
    with pytest.raises(AssertionError):
        ao.process_adjacency_lists([np.zeros((1, 2), np.int32)], 2, True, set(), 4)
    rc, _ = _sizes([1], 2, True, set(), 4)
    assert rc == _ffi.ERR_INVALID_ARGUMENT
    assert b"self_loop_edge_type" in _ffi.lib().tfgnn_b200_last_error()


def test_tie_and_count_helpers_match_oracle():
    for tie in (True, False, [0, 2], []):
        for T in (1, 3):
            assert get_tied_edge_types(tie, T) == ao.get_tied_edge_types(tie, T)
            for self_loops in (True, False):
                tied = get_tied_edge_types(tie, T)
                assert (compute_number_of_edge_types(tied, T, self_loops)
                        == ao.compute_number_of_edge_types(tied, T, self_loops))


def _reference_rule(node_counts, max_nodes):
    """graph_dataset.py:164-188 restated with explicit batch state."""
    batches, cur, nodes = [], [], 0
    for g, n in enumerate(node_counts):
        if nodes + n > max_nodes:
            batches.append(cur)
            cur, nodes = [], 0
        cur.append(g)
        nodes += n
    batches.append(cur)
    return [b for b in batches if b]


@pytest.mark.parametrize("max_nodes", [10, 37, 1000])
def test_greedy_batching_rule(max_nodes):
    rng = np.random.default_rng(max_nodes)
    counts = rng.integers(1, 30, size=200)
    got = [b.tolist() for b in greedy_batches(counts, max_nodes)]
    assert got == _reference_rule(counts.tolist(), max_nodes)
    assert sum(len(b) for b in got) == 200
    for b in got:
        assert counts[b].sum() <= max_nodes or len(b) == 1


def test_workspace_bytes():
    assert _ffi.lib().tfgnn_b200_assemble_batch_workspace_bytes(3, 10) == 4 * 11 * 8


def test_greedy_batching_rule_matches_executed_reference():
    """The partition of graphs into batches produced by the reference's own iterator (batch_assembly_golden.json)."""
    path = os.path.join(os.path.dirname(GOLDEN), "batch_assembly_golden.json")
    with open(path) as f:
        cases = json.load(f)["cases"]
    for c in cases:
        counts = [len(g["node_features"]) for g in c["graphs"]]
        got = [b.tolist() for b in greedy_batches(counts, c["max_nodes_per_batch"])]
        assert got == [b["graph_ids"] for b in c["batches"]]
