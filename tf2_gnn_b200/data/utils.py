"""tf2_gnn/data/utils.py, same names and argument meaning, arrays on the GPU.

``process_adjacency_lists`` runs ``tfgnn_b200_process_adjacency`` (batch_builder.cu): the flips, the tied / fresh
backward types, the self-loop type and the in-degree table are built by CUDA kernels from device-resident int32
edge lists; nothing is computed on the host.  Results are bit-identical to the reference's numpy code
(tests/golden/process_adjacency_lists_golden.json).
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_int32, c_int64, c_void_p
from typing import List, Sequence, Set, Tuple, Union

import torch

from .. import _ffi
from ..runtime import require_cuda, stream_ptr, to_device_adj


def get_tied_edge_types(tie_fwd_bkwd_edges: Union[bool, List[int]], num_fwd_edge_types: int) -> Set[int]:
    """data/utils.py:61-77: a list names the tied forward types, True ties all, False none."""
    if isinstance(tie_fwd_bkwd_edges, list):
        return set(tie_fwd_bkwd_edges)
    return set(range(num_fwd_edge_types)) if tie_fwd_bkwd_edges else set()


def compute_number_of_edge_types(tied_fwd_bkwd_edge_types: Set[int], num_fwd_edge_types: int,
                                 add_self_loop_edges: bool) -> int:
    """data/utils.py:80-84."""
    return 2 * num_fwd_edge_types - len(tied_fwd_bkwd_edge_types) + int(add_self_loop_edges)


def process_adjacency_lists(adjacency_lists: Sequence, num_nodes: int, add_self_loop_edges: bool,
                            tied_fwd_bkwd_edge_types: Set[int], self_loop_edge_type: int = 0
                            ) -> Tuple[List[torch.Tensor], torch.Tensor]:
    """data/utils.py:9-58 on the device.

    Returns (processed adjacency lists: int32 CUDA tensors [E_l, 2], type_to_num_incoming_edges: float32 CUDA tensor
    [L, num_nodes]).  The reference returns numpy arrays and float64 counts of the same integer values.
    """
    dev = require_cuda()
    fwd = [to_device_adj(a, dev) for a in adjacency_lists]
    T = len(fwd)
    E = (c_int64 * max(T, 1))(*[int(a.shape[0]) for a in fwd])
    tied = (c_int32 * max(T, 1))(*[1 if t in tied_fwd_bkwd_edge_types else 0 for t in range(T)])
    L = c_int32(0)
    E_out = (c_int64 * (2 * T + 1))()
    lib = _ffi.lib()
    rc = lib.tfgnn_b200_process_adjacency_sizes(E, T, int(num_nodes), int(bool(add_self_loop_edges)), tied,
                                                int(self_loop_edge_type), E_out, byref(L))
    if rc == _ffi.ERR_INVALID_ARGUMENT and b"self_loop_edge_type" in lib.tfgnn_b200_last_error():
        n_types = 2 * T - len([t for t in range(T) if t in tied_fwd_bkwd_edge_types])
        raise AssertionError(   # the reference asserts (data/utils.py:93-97)
            f"Self loop edge type {self_loop_edge_type} should be in range [{-(n_types + 1)}, {n_types}].")
    _ffi.check(rc)
    out = [torch.empty((int(E_out[l]), 2), dtype=torch.int32, device=dev) for l in range(L.value)]
    counts = torch.empty((L.value, int(num_nodes)), dtype=torch.float32, device=dev)
    in_ptrs = (c_void_p * max(T, 1))(*[a.data_ptr() if a.numel() else None for a in fwd])
    out_ptrs = (c_void_p * max(L.value, 1))(*[a.data_ptr() if a.numel() else None for a in out])
    _ffi.check(lib.tfgnn_b200_process_adjacency(
        ctypes.cast(in_ptrs, _ffi._PP), E, T, int(num_nodes), int(bool(add_self_loop_edges)), tied,
        int(self_loop_edge_type), ctypes.cast(out_ptrs, _ffi._PP), L.value,
        counts.data_ptr() if counts.numel() else None, stream_ptr()))
    return out, counts
