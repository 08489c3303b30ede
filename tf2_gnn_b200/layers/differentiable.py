"""Differentiable generic message passing.

The reference trains EVERY variant by letting tf.GradientTape differentiate its literal op sequence
(message_passing.py:95-218: gather -> _message_function -> concat -> optional activation -> unsorted_segment_* ->
activation).  The fused forward kernels reorder that sequence; RGCN-style layers and GGNN have fused backward kernels
(csrc/backward.cu).  Every other configuration of the Edge-MLP family — hidden layers in the edge MLPs, RGIN (with its
aggregation MLP), GNN-FiLM, max aggregation, activation before aggregation — trains through THIS module: the
reference's own op order, each op a C-ABI kernel with a C-ABI backward (gather_rows <-> unsorted_segment_sum are each
other's adjoint).  It materialises [E, D] tensors exactly like the reference does; it is the correctness path for
training, not the fast path for inference (inference never comes here).

torch is used as the autograd tape and for data movement (cat / slicing) only.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from .. import _ffi
from ..runtime import PreparedBatch, stream_ptr
from . import node_ops


def _f32(shape, like):
    return torch.empty(shape, dtype=torch.float32, device=like.device)


# ---- gather / segment reduce (adjoint pair) ---------------------------------------------------------------------
def _gather_rows(table: torch.Tensor, adj: torch.Tensor, column: int) -> torch.Tensor:
    """table[adj[:, column]] — tf.nn.embedding_lookup (message_passing.py:197-206)."""
    E, D = int(adj.shape[0]), int(table.shape[1])
    out = _f32((E, D), table)
    if E:
        _ffi.check(_ffi.lib().tfgnn_b200_gather_rows(table.data_ptr(), int(table.shape[0]), D, adj.data_ptr() + 4 * column, 2,
                                                     E, out.data_ptr(), stream_ptr()))
    return out


def _segment_sum_by(adj: torch.Tensor, column: int, data: torch.Tensor, num_segments: int) -> torch.Tensor:
    out = _f32((num_segments, int(data.shape[1])), data)
    _ffi.check(_ffi.lib().tfgnn_b200_unsorted_segment_reduce(
        data.data_ptr(), adj.data_ptr() + 4 * column, 2, int(data.shape[0]), int(data.shape[1]), num_segments,
        _ffi.AGG["sum"], out.data_ptr(), stream_ptr()))
    return out


class _GatherFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, adj, column):
        ctx.adj, ctx.column, ctx.V = adj, column, int(table.shape[0])
        return _gather_rows(table.contiguous(), adj, column)

    @staticmethod
    def backward(ctx, g):
        return _segment_sum_by(ctx.adj, ctx.column, g.contiguous(), ctx.V), None, None


def gather(table: torch.Tensor, adj: torch.Tensor, column: int) -> torch.Tensor:
    if node_ops._needs_grad(table):
        return _GatherFunction.apply(table, adj, column)
    return _gather_rows(table.contiguous(), adj, column)


class _SegmentReduceFunction(torch.autograd.Function):
    """tf.math.unsorted_segment_{sum,mean,sqrt_n,max} (message_passing.py:172-174) keyed by a 1-D int32 id tensor."""

    @staticmethod
    def forward(ctx, data, ids, num_segments, agg):
        data = data.contiguous()
        M, H = int(data.shape[0]), int(data.shape[1])
        out = _f32((num_segments, H), data)
        _ffi.check(_ffi.lib().tfgnn_b200_unsorted_segment_reduce(data.data_ptr(), ids.data_ptr(), 1, M, H, num_segments,
                                                                 _ffi.AGG[agg], out.data_ptr(), stream_ptr()))
        ctx.ids, ctx.agg, ctx.num_segments = ids, agg, num_segments
        ctx.save_for_backward(data, out)
        return out

    @staticmethod
    def backward(ctx, g):
        data, out = ctx.saved_tensors
        g = g.contiguous()
        M, H = int(data.shape[0]), int(data.shape[1])
        lib = _ffi.lib()
        grad = _f32((M, H), data)
        if M == 0:
            return grad, None, None, None
        if ctx.agg == "max":
            _ffi.check(lib.tfgnn_b200_segment_max_bwd(data.data_ptr(), ctx.ids.data_ptr(), 1, out.data_ptr(), g.data_ptr(), M, H,
                                                      ctx.num_segments, grad.data_ptr(), stream_ptr()))
            return grad, None, None, None
        _ffi.check(lib.tfgnn_b200_gather_rows(g.data_ptr(), ctx.num_segments, H, ctx.ids.data_ptr(), 1, M, grad.data_ptr(),
                                              stream_ptr()))
        if ctx.agg in ("mean", "sqrt_n"):
            ones = torch.ones((M, 1), dtype=torch.float32, device=data.device)
            counts = _f32((ctx.num_segments, 1), data)
            _ffi.check(lib.tfgnn_b200_unsorted_segment_reduce(ones.data_ptr(), ctx.ids.data_ptr(), 1, M, 1, ctx.num_segments,
                                                              _ffi.AGG["sum"], counts.data_ptr(), stream_ptr()))
            per_msg = _f32((M, 1), data)
            _ffi.check(lib.tfgnn_b200_gather_rows(counts.data_ptr(), ctx.num_segments, 1, ctx.ids.data_ptr(), 1, M,
                                                  per_msg.data_ptr(), stream_ptr()))
            scaled = _f32((M, H), data)
            _ffi.check(lib.tfgnn_b200_row_scale(grad.data_ptr(), per_msg.data_ptr(), M, H, 2 if ctx.agg == "mean" else 3,
                                                scaled.data_ptr(), stream_ptr()))
            grad = scaled
        return grad, None, None, None


# ---- elementwise ------------------------------------------------------------------------------------------------
class _ActivationFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, code):
        x = x.contiguous()
        out = torch.empty_like(x)
        _ffi.check(_ffi.lib().tfgnn_b200_activation(x.data_ptr(), x.numel(), code, out.data_ptr(), stream_ptr()))
        ctx.code = code
        ctx.save_for_backward(x if code == _ffi.ACT["gelu"] else out)
        return out

    @staticmethod
    def backward(ctx, g):
        (ref,) = ctx.saved_tensors
        g = g.contiguous()
        out = torch.empty_like(g)
        _ffi.check(_ffi.lib().tfgnn_b200_activation_bwd(ref.data_ptr(), g.data_ptr(), g.numel(), ctx.code, out.data_ptr(),
                                                        stream_ptr()))
        return out, None


def activation(x: torch.Tensor, act) -> torch.Tensor:
    if act is None:
        return x
    if node_ops._needs_grad(x):
        return _ActivationFunction.apply(x, act.code)
    return act(x)


class _RowScaleFunction(torch.autograd.Function):
    """messages * 1/(n + 1e-7) (gnn_edge_mlp.py:102-106); the scale carries no gradient (it is a count)."""

    @staticmethod
    def forward(ctx, x, s, mode):
        x = x.contiguous()
        out = torch.empty_like(x)
        if x.shape[0]:
            _ffi.check(_ffi.lib().tfgnn_b200_row_scale(x.data_ptr(), s.data_ptr(), int(x.shape[0]), int(x.shape[1]), mode,
                                                       out.data_ptr(), stream_ptr()))
        ctx.s, ctx.mode = s, mode
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        out = torch.empty_like(g)
        if g.shape[0]:
            _ffi.check(_ffi.lib().tfgnn_b200_row_scale(g.data_ptr(), ctx.s.data_ptr(), int(g.shape[0]), int(g.shape[1]),
                                                       ctx.mode, out.data_ptr(), stream_ptr()))
        return out, None, None


def _mul_add(a, lda, b, ldb, c, ldc, M, H, out, ldo):
    if M:
        _ffi.check(_ffi.lib().tfgnn_b200_mul_add(a, lda, b, ldb, c, ldc, M, H, out, ldo, stream_ptr()))


class _FilmFunction(torch.autograd.Function):
    """film[:, :H] * messages + film[:, H:] (gnn_film.py:103-107)."""

    @staticmethod
    def forward(ctx, film, messages):
        film, messages = film.contiguous(), messages.contiguous()
        E, H = int(messages.shape[0]), int(messages.shape[1])
        out = torch.empty_like(messages)
        _mul_add(film.data_ptr(), 2 * H, messages.data_ptr(), H, film.data_ptr() + 4 * H, 2 * H, E, H, out.data_ptr(), H)
        ctx.save_for_backward(film, messages)
        return out

    @staticmethod
    def backward(ctx, g):
        film, messages = ctx.saved_tensors
        g = g.contiguous()
        E, H = int(messages.shape[0]), int(messages.shape[1])
        g_film = torch.empty_like(film)
        g_msg = torch.empty_like(messages)
        _mul_add(g.data_ptr(), H, messages.data_ptr(), H, 0, 0, E, H, g_film.data_ptr(), 2 * H)       # d/dgamma = g * m
        if E:
            g_film[:, H:] = g                                                                      # d/dbeta  = g (copy)
        _mul_add(g.data_ptr(), H, film.data_ptr(), 2 * H, 0, 0, E, H, g_msg.data_ptr(), H)           # d/dm     = g * gamma
        return g_film, g_msg


# ---- the literal layer ------------------------------------------------------------------------------------------
def _mlp(x: torch.Tensor, kernels: Sequence[torch.Tensor]) -> torch.Tensor:
    """dpu_utils MLP of the message functions: bias-free, ReLU hidden layers, linear output."""
    from ..utils.param_helpers import get_activation_function
    relu = get_activation_function("relu")
    cur = x
    for i, W in enumerate(kernels):
        cur = node_ops.dense(cur, W, None, relu if i < len(kernels) - 1 else None)
    return cur


def edge_mlp_family_forward(layer, h: torch.Tensor, prepared: PreparedBatch, *, film_kernels: Optional[List] = None,
                            aggr_kernels: Optional[List[torch.Tensor]] = None, final_activation: bool = True,
                            activation_before: Optional[bool] = None) -> torch.Tensor:
    """message_passing.py:95-218 with GNN_Edge_MLP._message_function (gnn_edge_mlp.py:84-107), optionally GNN_FiLM's
    modulation (gnn_film.py:83-108) and RGIN's aggregation MLP (rgin.py:88-106), in the reference's op order."""
    if prepared.target_range != (0, prepared.num_source_nodes):
        raise NotImplementedError("training through a target-range shard is not built")
    V = int(h.shape[0])
    in_degree = prepared.in_degree()                                   # [L, V] (message_passing.py:190)
    messages_per_type, targets = [], []
    for l, adj in enumerate(prepared.adjacency_lists):
        E = int(adj.shape[0])
        src = gather(h, adj, 0)                                         # :197-200
        tgt = gather(h, adj, 1) if (layer._use_target_state_as_input or film_kernels is not None) else None   # :201-204
        x = torch.cat([src, tgt], dim=1) if layer._use_target_state_as_input else src
        m = _mlp(x, [v.value for v in layer._edge_type_mlps[l].layers])  # gnn_edge_mlp.py:100
        if layer._normalize_by_num_incoming:                            # :102-106
            n_in = _gather_rows(in_degree[l].reshape(V, 1).contiguous(), adj, 1)
            m = _RowScaleFunction.apply(m, n_in, 1)
        if film_kernels is not None:                                    # gnn_film.py:99-107
            film = _mlp(tgt, film_kernels[l])
            m = _FilmFunction.apply(film, m)
        messages_per_type.append(m)
        targets.append(adj[:, 1])
    H = layer._hidden_dim
    if messages_per_type:
        messages = torch.cat(messages_per_type, dim=0)                  # message_passing.py:166-167
        ids = torch.cat([t.reshape(-1) for t in targets], dim=0).contiguous()
    else:
        messages = torch.zeros((0, H), dtype=torch.float32, device=h.device)
        ids = torch.zeros((0,), dtype=torch.int32, device=h.device)
    before = layer._message_activation_before_aggregation if activation_before is None else activation_before
    if before:
        messages = activation(messages, layer._activation_fn)           # :169-170
    agg = _SegmentReduceFunction.apply(messages, ids, V, layer._aggregation_fn.name)   # :172-174
    if aggr_kernels is not None:
        agg = _mlp(agg, aggr_kernels)                                   # rgin.py:103-104
    if final_activation and not before:
        agg = activation(agg, layer._activation_fn)                     # :176-177
    return agg

