"""TensorFlow side of the drop-in boundary: zero-copy DLPack bridge onto the C ABI.

north_star: "Host code stays Python/TF2, calling hand-written sm_100a CUDA kernels through a thin
C-ABI/ctypes (dlpack zero-copy from TF tensors)".  This module is that bridge.  It has two layers:

* ``DLTensorView`` / ``dlpack_view`` — a pure-ctypes reader of a DLPack capsule (the ``dltensor``
  PyCapsule every DLPack producer hands out: ``tf.experimental.dlpack.to_dlpack``,
  ``torch.utils.dlpack.to_dlpack``, ``array.__dlpack__()``).  It needs neither TensorFlow nor torch
  and is unit-tested on CPU with numpy / torch capsules (tests/test_tf_adapter_cpu.py).
* ``TFBackend`` — the calls a reference maintainer makes from
  ``tf2_gnn/layers/message_passing/*.py``: ``prepare`` once per batch
  (gnn.py:278,301 hands the same adjacency lists to every layer) and one ``*_forward`` per layer
  (message_passing.py:95-133).  TensorFlow is imported lazily; it is absent from this build image, so
  this half runs wherever TF exists (the shipped, GPU-tested carrier is the torch mirror in
  ``tf2_gnn_b200.layers``).

Nothing here computes: pointers in, pointers out.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, Structure, byref, c_int32, c_int64, c_uint8, c_uint16, c_uint32, c_uint64, c_void_p
from typing import Any, List, Optional, Sequence, Tuple

from . import _ffi

# ---- DLPack ABI (dlpack.h, v0.x unversioned capsule "dltensor") -------------------------------
kDLCPU, kDLCUDA, kDLCUDAHost, kDLCUDAManaged = 1, 2, 3, 13
kDLInt, kDLUInt, kDLFloat, kDLBfloat = 0, 1, 2, 4


class DLDevice(Structure):
    _fields_ = [("device_type", c_int32), ("device_id", c_int32)]


class DLDataType(Structure):
    _fields_ = [("code", c_uint8), ("bits", c_uint8), ("lanes", c_uint16)]


class DLTensor(Structure):
    _fields_ = [("data", c_void_p), ("device", DLDevice), ("ndim", c_int32), ("dtype", DLDataType),
                ("shape", POINTER(c_int64)), ("strides", POINTER(c_int64)), ("byte_offset", c_uint64)]


class DLManagedTensor(Structure):
    pass


DLManagedTensor._fields_ = [("dl_tensor", DLTensor), ("manager_ctx", c_void_p),
                            ("deleter", ctypes.CFUNCTYPE(None, POINTER(DLManagedTensor)))]

_PyCapsule_GetPointer = ctypes.pythonapi.PyCapsule_GetPointer
_PyCapsule_GetPointer.restype = c_void_p
_PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]
_PyCapsule_IsValid = ctypes.pythonapi.PyCapsule_IsValid
_PyCapsule_IsValid.restype = ctypes.c_int
_PyCapsule_IsValid.argtypes = [ctypes.py_object, ctypes.c_char_p]
_PyCapsule_SetName = ctypes.pythonapi.PyCapsule_SetName
_PyCapsule_SetName.restype = ctypes.c_int
_PyCapsule_SetName.argtypes = [ctypes.py_object, ctypes.c_char_p]


class DLTensorView:
    """What the C ABI needs from a tensor: address, shape, dtype, device.  Holds the capsule (and through
    it the producer's buffer) alive; ``release()`` consumes the capsule the way a DLPack consumer must
    (rename to ``used_dltensor`` and call the deleter), ``__del__`` does it if the caller forgot."""

    def __init__(self, capsule: Any):
        if not _PyCapsule_IsValid(capsule, b"dltensor"):
            raise ValueError("not a live DLPack capsule (expected a PyCapsule named 'dltensor')")
        self._capsule = capsule
        self._managed = ctypes.cast(_PyCapsule_GetPointer(capsule, b"dltensor"), POINTER(DLManagedTensor))
        t = self._managed.contents.dl_tensor
        self.ndim = int(t.ndim)
        self.shape: Tuple[int, ...] = tuple(int(t.shape[i]) for i in range(self.ndim))
        self.strides: Optional[Tuple[int, ...]] = (
            tuple(int(t.strides[i]) for i in range(self.ndim)) if t.strides else None)
        self.dtype_code, self.dtype_bits, self.dtype_lanes = int(t.dtype.code), int(t.dtype.bits), int(t.dtype.lanes)
        self.device_type, self.device_id = int(t.device.device_type), int(t.device.device_id)
        self.data_ptr = int(t.data or 0) + int(t.byte_offset)
        self._released = False

    # -- checks the boundary relies on (include/tfgnn_b200.h "Conventions") ---------------------
    def is_contiguous(self) -> bool:
        if self.strides is None:
            return True
        expect = 1
        for dim, stride in zip(reversed(self.shape), reversed(self.strides)):
            if dim != 1 and stride != expect:
                return False
            expect *= dim
        return True

    def require(self, code: int, bits: int, ndim: Optional[int] = None, on_cuda: bool = True) -> "DLTensorView":
        """ValueError like the reference raises on a malformed input (never a silent conversion)."""
        if (self.dtype_code, self.dtype_bits, self.dtype_lanes) != (code, bits, 1):
            raise ValueError(f"expected dtype code {code}/{bits} bits, got {self.dtype_code}/{self.dtype_bits}")
        if ndim is not None and self.ndim != ndim:
            raise ValueError(f"expected a rank-{ndim} tensor, got shape {self.shape}")
        if not self.is_contiguous():
            raise ValueError("tensor must be C-contiguous (row-major)")
        if on_cuda and self.device_type not in (kDLCUDA, kDLCUDAManaged):
            raise ValueError("tensor must live in CUDA device memory (no CPU fallback exists)")
        return self

    def release(self) -> None:
        if self._released:
            return
        self._released = True
        deleter = self._managed.contents.deleter
        _PyCapsule_SetName(self._capsule, b"used_dltensor")
        if deleter:
            deleter(self._managed)

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def dlpack_view(x: Any) -> DLTensorView:
    """View of anything that speaks DLPack: a raw capsule, an object with ``__dlpack__`` (numpy, torch, cupy,
    jax), or a ``tf.Tensor`` (through ``tf.experimental.dlpack.to_dlpack``)."""
    if type(x).__name__ == "PyCapsule":
        return DLTensorView(x)
    mod = type(x).__module__ or ""
    if mod.startswith("tensorflow"):
        import tensorflow as tf  # lazy: absent from the build image
        return DLTensorView(tf.experimental.dlpack.to_dlpack(x))
    if hasattr(x, "__dlpack__"):
        return DLTensorView(x.__dlpack__())
    raise TypeError(f"cannot take a DLPack view of {type(x)!r}")


# ------------------------------------------------------------------------------------------------
# The calls the reference's layers make
# ------------------------------------------------------------------------------------------------
class TFPreparedBatch:
    """tfgnn_b200_prepare on TF adjacency lists, once per batch (replaces the per-layer
    calculate_type_to_num_incoming_edges, message_passing.py:190,230-263)."""

    def __init__(self, adjacency_lists: Sequence[Any], num_nodes: int, stream: Optional[int] = None):
        self.handle = c_void_p()
        self._views: List[DLTensorView] = []
        self._views = [dlpack_view(a).require(kDLInt, 32, ndim=2) for a in adjacency_lists]
        for v in self._views:
            if v.shape[1] != 2:
                raise ValueError("adjacency lists must have shape [E, 2]")
        L = len(self._views)
        if L > _ffi.MAX_EDGE_TYPES:
            raise ValueError(f"at most {_ffi.MAX_EDGE_TYPES} edge types are supported")
        ptrs = (c_void_p * max(L, 1))(*[v.data_ptr for v in self._views])
        counts = (c_int64 * max(L, 1))(*[v.shape[0] for v in self._views])
        self.num_nodes = int(num_nodes)
        _ffi.check(_ffi.lib().tfgnn_b200_prepare(ptrs, counts, L, self.num_nodes, 0, byref(self.handle),
                                                 c_void_p(stream or 0)))

    def close(self) -> None:
        if self.handle:
            _ffi.lib().tfgnn_b200_free_batch(self.handle)
            self.handle = c_void_p()
        for v in self._views:
            v.release()
        self._views = []

    __del__ = close


class TFBackend:
    """Per-layer forward calls on TF tensors.  Outputs are allocated by TensorFlow (``tf.zeros``: caller-owned,
    pre-allocated, as the ABI requires) and written in place through their DLPack pointer.

    Stream: TensorFlow does not expose its compute stream to Python.  ``stream=None`` enqueues on the legacy
    default stream, which synchronises with TF's blocking streams; a custom-op wrapper passes
    ``ctx->eigen_device<GPUDevice>().stream()`` instead."""

    def __init__(self, stream: Optional[int] = None):
        self.stream = c_void_p(stream or 0)

    @staticmethod
    def _tf():
        try:
            import tensorflow as tf
        except ImportError as e:  # pragma: no cover - TF is absent from the build image
            raise ImportError("tf_adapter.TFBackend needs tensorflow>=2.0 (the torch carrier in "
                              "tf2_gnn_b200.layers is the one shipped with this image)") from e
        return tf

    def prepare(self, adjacency_lists: Sequence[Any], num_nodes: int) -> TFPreparedBatch:
        return TFPreparedBatch(adjacency_lists, num_nodes, stream=self.stream.value)

    def _weights(self, kernels: Sequence[Any]) -> Tuple[Any, List[DLTensorView]]:
        views = [dlpack_view(k).require(kDLFloat, 32, ndim=2) for k in kernels]
        return (c_void_p * max(len(views), 1))(*[v.data_ptr for v in views]), views

    def edge_mlp_forward(self, prepared: TFPreparedBatch, node_embeddings: Any, mlp_kernels: Sequence[Any],
                         num_hidden_layers: int, hidden_dim: int, *, normalize: bool, use_target_state: bool,
                         act_before_aggregation: bool = False, aggregation: str = "sum",
                         activation: Optional[str] = "relu", path: str = "auto"):
        """GNN_Edge_MLP / RGCN layer (gnn_edge_mlp.py:84-107 + message_passing.py:95-179).  `mlp_kernels` is the
        type-major flat list of the bias-free Dense kernels of `self._edge_type_mlps`."""
        tf = self._tf()
        h = dlpack_view(node_embeddings).require(kDLFloat, 32, ndim=2)
        wptrs, wviews = self._weights(mlp_kernels)
        out = tf.zeros([h.shape[0], hidden_dim], tf.float32)
        o = dlpack_view(out).require(kDLFloat, 32, ndim=2)
        flags = ((_ffi.FLAG_NORMALIZE if normalize else 0) | (_ffi.FLAG_USE_TARGET if use_target_state else 0)
                 | (_ffi.FLAG_ACT_BEFORE_AGG if act_before_aggregation else 0))
        try:
            _ffi.check(_ffi.lib().tfgnn_b200_edge_mlp_fwd(
                prepared.handle, c_void_p(h.data_ptr), c_int32(h.shape[1]), wptrs, c_int32(num_hidden_layers),
                c_int32(hidden_dim), c_uint32(flags), c_int32(_ffi.AGG[aggregation]), c_int32(_ffi.ACT[activation]),
                c_int32(_ffi.PATH[path]), c_void_p(o.data_ptr), self.stream))
        finally:
            for v in [h, o, *wviews]:
                v.release()
        return out

    def rgcn_forward(self, prepared: TFPreparedBatch, node_embeddings: Any, kernels: Sequence[Any], hidden_dim: int,
                     normalize: bool = True, aggregation: str = "sum", activation: Optional[str] = "relu"):
        """RGCN.call (rgcn.py:12-62)."""
        return self.edge_mlp_forward(prepared, node_embeddings, kernels, 0, hidden_dim, normalize=normalize,
                                     use_target_state=False, aggregation=aggregation, activation=activation)
