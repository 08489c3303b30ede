"""The DLPack half of tf_adapter.py on CPU: capsule -> (pointer, shape, dtype, device), the checks the boundary
makes, and capsule consumption.  Producers here are numpy and torch (TensorFlow is absent from the image);
tf.experimental.dlpack.to_dlpack hands out the same 'dltensor' capsule."""
import ctypes
import gc

import numpy as np
import pytest

from tf2_gnn_b200 import tf_adapter as ta


def test_view_of_a_numpy_capsule_reads_pointer_shape_dtype():
    a = np.arange(24, dtype=np.float32).reshape(6, 4)
    v = ta.dlpack_view(a)
    assert v.shape == (6, 4) and v.ndim == 2
    assert (v.dtype_code, v.dtype_bits, v.dtype_lanes) == (ta.kDLFloat, 32, 1)
    assert v.device_type == ta.kDLCPU
    assert v.data_ptr == a.ctypes.data
    # zero-copy: read the producer's memory through the pointer
    got = np.ctypeslib.as_array(ctypes.cast(v.data_ptr, ctypes.POINTER(ctypes.c_float)), shape=(24,))
    assert np.array_equal(got, a.reshape(-1))
    v.release()


def test_view_of_a_torch_capsule_and_int32_adjacency():
    torch = pytest.importorskip("torch")
    adj = torch.tensor([[0, 1], [2, 3], [4, 0]], dtype=torch.int32)
    v = ta.dlpack_view(torch.utils.dlpack.to_dlpack(adj))
    assert v.shape == (3, 2) and (v.dtype_code, v.dtype_bits) == (ta.kDLInt, 32)
    assert v.data_ptr == adj.data_ptr()
    v.require(ta.kDLInt, 32, ndim=2, on_cuda=False)
    with pytest.raises(ValueError):
        v.require(ta.kDLFloat, 32, on_cuda=False)      # wrong dtype: never a silent conversion
    with pytest.raises(ValueError):
        v.require(ta.kDLInt, 32, ndim=2, on_cuda=True)  # host memory: there is no CPU fallback
    v.release()


def test_byte_offset_and_non_contiguous_views():
    a = np.arange(40, dtype=np.float32).reshape(10, 4)
    v = ta.dlpack_view(a[2:])                 # numpy exports the offset either in data or in byte_offset
    assert v.data_ptr == a[2:].ctypes.data and v.is_contiguous()
    v.release()
    t = ta.dlpack_view(a[:, ::2])
    assert not t.is_contiguous()
    with pytest.raises(ValueError):
        t.require(ta.kDLFloat, 32, ndim=2, on_cuda=False)
    t.release()


def test_capsule_is_consumed_exactly_once():
    torch = pytest.importorskip("torch")
    x = torch.ones(8)
    cap = torch.utils.dlpack.to_dlpack(x)
    v = ta.DLTensorView(cap)
    v.release()
    v.release()                               # idempotent
    with pytest.raises(ValueError):
        ta.DLTensorView(cap)                  # renamed to 'used_dltensor': cannot be consumed twice
    del v
    gc.collect()


def test_prepare_validates_adjacency_shape_before_any_cuda_call():
    bad = np.zeros((4, 3), dtype=np.int32)
    with pytest.raises(ValueError):
        ta.TFPreparedBatch([bad], 5)          # host tensor / wrong shape -> ValueError, like the reference's asserts


def test_backend_without_tensorflow_raises_importerror():
    try:
        import tensorflow  # noqa: F401
        pytest.skip("tensorflow is installed")
    except ImportError:
        pass
    with pytest.raises(ImportError):
        ta.TFBackend._tf()
