"""Node-level and graph-level ops of GNN._internal_call (gnn.py:279-327) on the library's kernels, each with
the gradient the reference obtains from tf.GradientTape (models/graph_task_model.py:338-365).

Every function takes and returns CUDA float32 tensors; when autograd is recording and an input requires grad the
op runs through a torch.autograd.Function whose backward is again a C-ABI call — no torch arithmetic on the path.
Ops whose backward is not built (the graph readout / exchange family) raise under autograd instead of silently
cutting the gradient ("training through the GNN stack silently truncates gradients").
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import _ffi
from ..runtime import stream_ptr


def _needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


def require_no_grad(what: str, *tensors) -> None:
    if _needs_grad(*tensors):
        raise NotImplementedError(
            f"{what}: the backward pass of this op is not built; run it under torch.no_grad() or detach its inputs "
            "(it never silently drops a gradient)")


# ---- Dense ------------------------------------------------------------------------------------------------------
def _dense_fwd(x, W, bias, act_code):
    V, K = int(x.shape[0]), int(x.shape[1])
    N = int(W.shape[1])
    out = torch.empty((V, N), dtype=torch.float32, device=x.device)
    if bias is None:
        _ffi.check(_ffi.lib().tfgnn_b200_dense_fwd(x.data_ptr(), W.data_ptr(), out.data_ptr(), V, K, N, act_code, 0,
                                                   stream_ptr()))
    else:
        _ffi.check(_ffi.lib().tfgnn_b200_dense_bias_fwd(x.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), V, K,
                                                        N, act_code, 0, stream_ptr()))
    return out


class _DenseFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, bias, act_code):
        out = _dense_fwd(x, W, bias, act_code)
        ctx.act_code = act_code
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, W, bias if bias is not None else W.new_empty(0), out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, W, bias, out = ctx.saved_tensors
        bias = bias if ctx.has_bias else None
        grad_out = grad_out.contiguous()
        V, K, N = int(x.shape[0]), int(x.shape[1]), int(W.shape[1])
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gW = torch.empty_like(W) if ctx.needs_input_grad[1] else None
        gb = torch.empty_like(bias) if (bias is not None and ctx.needs_input_grad[2]) else None
        _ffi.check(_ffi.lib().tfgnn_b200_dense_bwd(
            x.data_ptr(), W.data_ptr(), _ptr(bias), out.data_ptr(), grad_out.data_ptr(), V, K, N, ctx.act_code,
            _ptr(gx), _ptr(gW), _ptr(gb), stream_ptr()))
        return gx, gW, gb, None


def dense(x: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor] = None, activation=None) -> torch.Tensor:
    """tf.keras.layers.Dense(units, use_bias=bias is not None, activation=activation)."""
    code = activation.code if activation is not None else 0
    x = x.contiguous()
    if _needs_grad(x, W, bias):
        return _DenseFunction.apply(x, W, bias, code)
    return _dense_fwd(x, W, bias, code)


def mlp(x: torch.Tensor, kernels, biases=None, hidden_activation=None, training: bool = False, dropout_rate: float = 0.0,
        rng: Optional["DropoutState"] = None) -> torch.Tensor:
    """dpu_utils.tf2utils.MLP: hidden Dense layers with `hidden_activation` (ReLU by default), linear output layer;
    under training, dropout on the input of every layer."""
    from ..utils.param_helpers import get_activation_function
    act = hidden_activation or get_activation_function("relu")
    cur = x
    n = len(kernels)
    for i, W in enumerate(kernels):
        if training and dropout_rate > 0.0:
            cur = dropout(cur, dropout_rate, rng)
        b = biases[i] if biases is not None else None
        cur = dense(cur, W, b, act if i < n - 1 else None)
    return cur


# ---- LayerNormalization -----------------------------------------------------------------------------------------
def _ln_fwd(x, gamma, beta, eps):
    out = torch.empty_like(x)
    _ffi.check(_ffi.lib().tfgnn_b200_layer_norm(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), int(x.shape[0]),
                                                int(x.shape[1]), eps, out.data_ptr(), stream_ptr()))
    return out


class _LayerNormFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        ctx.eps = eps
        ctx.save_for_backward(x, gamma)
        return _ln_fwd(x, gamma, beta, eps)

    @staticmethod
    def backward(ctx, grad_out):
        x, gamma = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gg, gb = torch.empty_like(gamma), torch.empty_like(gamma)
        _ffi.check(_ffi.lib().tfgnn_b200_layer_norm_bwd(x.data_ptr(), gamma.data_ptr(), grad_out.data_ptr(), int(x.shape[0]),
                                                        int(x.shape[1]), ctx.eps, _ptr(gx), gg.data_ptr(), gb.data_ptr(),
                                                        stream_ptr()))
        return gx, gg, gb, None


def layer_norm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, epsilon: float = 1e-3) -> torch.Tensor:
    x = x.contiguous()
    if _needs_grad(x, gamma, beta):
        return _LayerNormFunction.apply(x, gamma, beta, float(epsilon))
    return _ln_fwd(x, gamma, beta, float(epsilon))


# ---- residual average, scaling ----------------------------------------------------------------------------------
def _axpby(a, alpha, b, beta):
    out = torch.empty_like(a)
    _ffi.check(_ffi.lib().tfgnn_b200_axpby(a.data_ptr(), float(alpha), _ptr(b), float(beta), a.numel(), out.data_ptr(),
                                           stream_ptr()))
    return out


class _AverageFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, last):
        out = torch.empty_like(x)
        _ffi.check(_ffi.lib().tfgnn_b200_residual_average(x.data_ptr(), last.data_ptr(), out.data_ptr(), x.numel(),
                                                          stream_ptr()))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        half = _axpby(grad_out.contiguous(), 0.5, None, 0.0)
        return half, half


def residual_average(x: torch.Tensor, last: torch.Tensor) -> torch.Tensor:
    """cur += last; cur /= 2 (gnn.py:294-295)."""
    x, last = x.contiguous(), last.contiguous()
    if _needs_grad(x, last):
        return _AverageFunction.apply(x, last)
    out = torch.empty_like(x)
    _ffi.check(_ffi.lib().tfgnn_b200_residual_average(x.data_ptr(), last.data_ptr(), out.data_ptr(), x.numel(),
                                                      stream_ptr()))
    return out


# ---- dropout ----------------------------------------------------------------------------------------------------
class DropoutState:
    """Seed + running offset of the Philox stream (one per model, like a tf.random.Generator): every dropout call
    consumes ceil(n / 4) counter values, so masks of successive calls are independent and a run is reproducible
    from its seed."""

    def __init__(self, seed: int = 0):
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.offset = 0

    def take(self, n: int) -> int:
        off = self.offset
        self.offset += (int(n) + 3) // 4
        return off


_default_dropout_state = DropoutState(0x5EED)


def _dropout_apply(x, rate, seed, offset):
    out = torch.empty_like(x)
    _ffi.check(_ffi.lib().tfgnn_b200_dropout(x.data_ptr(), x.numel(), float(rate), seed, offset, out.data_ptr(),
                                             stream_ptr()))
    return out


class _DropoutFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rate, seed, offset):
        ctx.cfg = (rate, seed, offset)
        return _dropout_apply(x, rate, seed, offset)

    @staticmethod
    def backward(ctx, grad_out):
        rate, seed, offset = ctx.cfg
        return _dropout_apply(grad_out.contiguous(), rate, seed, offset), None, None, None


def dropout(x: torch.Tensor, rate: float, state: Optional[DropoutState] = None) -> torch.Tensor:
    """tf.nn.dropout(x, rate) (gnn.py:285-289).  rate == 0 is the identity, as in TensorFlow."""
    if rate <= 0.0:
        return x
    state = state or _default_dropout_state
    x = x.contiguous()
    offset = state.take(x.numel())
    if _needs_grad(x):
        return _DropoutFunction.apply(x, float(rate), state.seed, offset)
    return _dropout_apply(x, float(rate), state.seed, offset)


# ---- graph-level primitives (forward only) ----------------------------------------------------------------------
def graph_offsets(node_to_graph_map: torch.Tensor, num_graphs: int, validate: bool = False) -> torch.Tensor:
    V = int(node_to_graph_map.shape[0])
    out = torch.empty(int(num_graphs) + 1, dtype=torch.int32, device=node_to_graph_map.device)
    _ffi.check(_ffi.lib().tfgnn_b200_graph_offsets(node_to_graph_map.data_ptr(), V, int(num_graphs), out.data_ptr(),
                                                   1 if validate else 0, stream_ptr()))
    return out


def segment_softmax(scores: torch.Tensor, graph_ptr: torch.Tensor) -> torch.Tensor:
    require_no_grad("segment_softmax", scores)
    scores = scores.contiguous()
    out = torch.empty_like(scores)
    _ffi.check(_ffi.lib().tfgnn_b200_segment_softmax(scores.data_ptr(), graph_ptr.data_ptr(), int(graph_ptr.shape[0]) - 1,
                                                     int(scores.shape[1]), out.data_ptr(), stream_ptr()))
    return out


def weighted_segment_sum(node_reprs: torch.Tensor, weights: Optional[torch.Tensor], graph_ptr: torch.Tensor,
                         num_heads: int, mean: bool = False) -> torch.Tensor:
    require_no_grad("weighted_segment_sum", node_reprs, weights)
    node_reprs = node_reprs.contiguous()
    G, GD = int(graph_ptr.shape[0]) - 1, int(node_reprs.shape[1])
    out = torch.empty((G, GD), dtype=torch.float32, device=node_reprs.device)
    _ffi.check(_ffi.lib().tfgnn_b200_weighted_segment_sum(node_reprs.data_ptr(), _ptr(weights), graph_ptr.data_ptr(), G, GD,
                                                          int(num_heads), 1 if mean else 0, out.data_ptr(), stream_ptr()))
    return out


def gathered_add(a: torch.Tensor, b: torch.Tensor, index: Optional[torch.Tensor], scale: float = 1.0,
                 activation=None) -> torch.Tensor:
    """act((a[v] + b[index[v]]) * scale)."""
    require_no_grad("gathered_add", a, b)
    a = a.contiguous()
    out = torch.empty_like(a)
    _ffi.check(_ffi.lib().tfgnn_b200_gathered_add(a.data_ptr(), b.data_ptr(), _ptr(index), int(a.shape[0]), int(a.shape[1]),
                                                  float(scale), activation.code if activation is not None else 0,
                                                  out.data_ptr(), stream_ptr()))
    return out


def gru_cell(inputs: torch.Tensor, inputs_row_index: Optional[torch.Tensor], state: torch.Tensor, kernel: torch.Tensor,
             recurrent_kernel: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """tf.keras.layers.GRUCell(units=H) (TF2 defaults: reset_after=True, bias [2,3H], gates z|r|h) applied to
    inputs[inputs_row_index[v]] with state[v].  The input half (inputs K + b0) is computed once per row of `inputs`
    (e.g. once per GRAPH for the global exchange) and picked up per node inside the gate kernel."""
    require_no_grad("gru_cell", inputs, state, kernel, recurrent_kernel, bias)
    H = int(state.shape[1])
    gx = dense(inputs, kernel, bias[0])
    gh = dense(state, recurrent_kernel, bias[1])
    state = state.contiguous()
    out = torch.empty_like(state)
    _ffi.check(_ffi.lib().tfgnn_b200_gru_gate_fwd(gx.data_ptr(), _ptr(inputs_row_index), gh.data_ptr(), state.data_ptr(),
                                                  int(state.shape[0]), H, out.data_ptr(), stream_ptr()))
    return out


def clamp_(x: torch.Tensor, lower: Optional[float], upper: Optional[float]) -> torch.Tensor:
    if lower is None and upper is None:
        return x
    _ffi.check(_ffi.lib().tfgnn_b200_clamp(x.data_ptr(), x.numel(), float(lower or 0.0), float(upper or 0.0),
                                           0 if lower is None else 1, 0 if upper is None else 1, stream_ptr()))
    return x
