"""Name -> op maps with the reference's names and error behaviour
(tf2_gnn/utils/param_helpers.py:7-42).  The returned objects carry the enum code the C ABI wants
and are also callable on device tensors (they launch the library's own kernels)."""
from __future__ import annotations

from typing import Optional

import torch

from .. import _ffi
from ..runtime import stream_ptr, to_device_f32


class AggregationFn:
    """unsorted_segment_{sum,max,mean,sqrt_n} stand-in: callable(data, segment_ids, num_segments)."""

    def __init__(self, name: str):
        self.name = name
        self.code = _ffi.AGG[name]

    def __call__(self, data, segment_ids, num_segments: int) -> torch.Tensor:
        data = to_device_f32(data)
        M = int(data.shape[0])
        data2 = data.reshape(M, -1)
        H = int(data2.shape[1]) if data2.dim() == 2 and data2.shape[1] else 1
        ids = segment_ids if isinstance(segment_ids, torch.Tensor) else torch.as_tensor(segment_ids)
        ids = ids.to(device=data.device, dtype=torch.int32)
        stride = int(ids.stride(0)) if ids.numel() else 1
        out = torch.empty((int(num_segments), H), dtype=torch.float32, device=data.device)
        _ffi.check(_ffi.lib().tfgnn_b200_unsorted_segment_reduce(
            data2.data_ptr(), ids.data_ptr(), stride, M, H, int(num_segments), self.code, out.data_ptr(),
            stream_ptr()))
        return out.reshape((int(num_segments),) + tuple(data.shape[1:]))

    def __repr__(self):
        return f"AggregationFn({self.name!r})"


class ActivationFn:
    def __init__(self, name: str):
        self.name = name
        self.code = _ffi.ACT[name]

    def __call__(self, x) -> torch.Tensor:
        x = to_device_f32(x)
        out = torch.empty_like(x)
        _ffi.check(_ffi.lib().tfgnn_b200_activation(x.data_ptr(), x.numel(), self.code, out.data_ptr(), stream_ptr()))
        return out

    def __repr__(self):
        return f"ActivationFn({self.name!r})"


def get_aggregation_function(aggregation_fn_name: str) -> AggregationFn:
    """tf2_gnn/utils/param_helpers.py:7-19."""
    if aggregation_fn_name not in _ffi.AGG:
        raise ValueError(f"Unknown aggregation function: {aggregation_fn_name}")
    return AggregationFn(aggregation_fn_name)


def get_activation_function(activation_fn_name: Optional[str]) -> Optional[ActivationFn]:
    """tf2_gnn/utils/param_helpers.py:22-42.  As in the reference, "linear" is in the table with
    value None and therefore raises "Unknown activation function"."""
    if activation_fn_name is None:
        return None
    activation_fn_name = activation_fn_name.lower()
    if activation_fn_name == "linear" or activation_fn_name not in _ffi.ACT:
        raise ValueError(f"Unknown activation function: {activation_fn_name}")
    return ActivationFn(activation_fn_name)
