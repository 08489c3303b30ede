"""GGNN — mirror of tf2_gnn/layers/message_passing/ggnn.py:12-89 on the B200 path."""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch

from ... import _ffi
from ...runtime import PreparedBatch, stream_ptr
from .gnn_edge_mlp import GNN_Edge_MLP
from .message_passing import MessagePassingInput, _last_dim, register_message_passing_implementation


class _GGNNFunction(torch.autograd.Function):
    """Autograd hook of the GGNN layer: forward = tfgnn_b200_ggnn_fwd, backward =
    tfgnn_b200_ggnn_bwd (everything but h is recomputed)."""

    @staticmethod
    def forward(ctx, h, prepared, cfg, gru_kernel, gru_recurrent_kernel, gru_bias, *weights):
        out = torch.empty((prepared.num_nodes, cfg["H"]), dtype=torch.float32, device=h.device)
        _ffi.check(_ffi.lib().tfgnn_b200_ggnn_fwd(
            prepared.handle, h.data_ptr(), int(h.shape[1]), _ffi.ptr_array(weights), cfg["n_hidden"], cfg["H"],
            cfg["flags"], cfg["agg"], gru_kernel.data_ptr(), gru_recurrent_kernel.data_ptr(), gru_bias.data_ptr(),
            cfg["path"], out.data_ptr(), stream_ptr()))
        ctx.prepared, ctx.cfg = prepared, cfg
        ctx.save_for_backward(h, gru_kernel, gru_recurrent_kernel, gru_bias, *weights)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        h, gru_kernel, gru_recurrent_kernel, gru_bias, *weights = ctx.saved_tensors
        cfg, prepared = ctx.cfg, ctx.prepared
        if cfg["n_hidden"] != 0:
            raise NotImplementedError("GGNN backward is built for message MLPs without hidden layers only")
        grad_out = grad_out.contiguous()
        grad_h = torch.empty_like(h)
        grad_w = [torch.empty_like(w) for w in weights]
        g_k, g_u, g_b = (torch.empty_like(t) for t in (gru_kernel, gru_recurrent_kernel, gru_bias))
        _ffi.check(_ffi.lib().tfgnn_b200_ggnn_bwd(
            prepared.handle, prepared.transposed().handle, h.data_ptr(), int(h.shape[1]), _ffi.ptr_array(weights),
            cfg["H"], cfg["flags"], cfg["agg"], gru_kernel.data_ptr(), gru_recurrent_kernel.data_ptr(),
            gru_bias.data_ptr(), grad_out.data_ptr(), grad_h.data_ptr(), _ffi.ptr_array(grad_w), g_k.data_ptr(),
            g_u.data_ptr(), g_b.data_ptr(), stream_ptr()))
        return (grad_h, None, None, g_k, g_u, g_b, *grad_w)


@register_message_passing_implementation
class GGNN(GNN_Edge_MLP):
    """h'_v = GRUCell(h_v, sum_l sum_{(u,v) in A_l} W_l h_u)  (ggnn.py:13-45).  The node embedding
    dimension must equal hidden_dim (ggnn.py:30).  GRU = Keras GRUCell(units=H) with TF2 defaults:
    kernel [D,3H], recurrent_kernel [H,3H], bias [2,3H], gates z|r|h, reset_after=True."""

    @classmethod
    def get_default_hyperparameters(cls):
        these_hypers = {
            "use_target_state_as_input": False,
            "normalize_by_num_incoming": True,
            "num_edge_MLP_hidden_layers": 0,
        }
        mp_hypers = super().get_default_hyperparameters()
        mp_hypers.update(these_hypers)
        return mp_hypers

    def __init__(self, params: Dict[str, Any], **kwargs):
        super().__init__(params, **kwargs)
        self._gru_kernel = None
        self._gru_recurrent_kernel = None
        self._gru_bias = None

    def build(self, input_shapes: MessagePassingInput):
        D = _last_dim(input_shapes.node_embeddings)
        H = self._hidden_dim
        self._gru_kernel = self.add_weight("gru_cell/kernel:0", (D, 3 * H))
        # Keras uses an orthogonal recurrent initialiser; Glorot keeps the same shape contract.
        self._gru_recurrent_kernel = self.add_weight("gru_cell/recurrent_kernel:0", (H, 3 * H))
        self._gru_bias = self.add_weight("gru_cell/bias:0", (2, 3 * H),
                                         initial_value=torch.zeros((2, 3 * H), dtype=torch.float32))
        super().build(input_shapes)

    def call(self, inputs: MessagePassingInput, training: bool = False,
             prepared: Optional[PreparedBatch] = None):
        h, prepared = self._device_inputs(inputs, prepared)
        self._check_types(prepared)
        if int(h.shape[1]) != self._hidden_dim:
            raise ValueError("GGNN: the node embedding dimension must equal hidden_dim")
        gru = (self._gru_kernel.value, self._gru_recurrent_kernel.value, self._gru_bias.value)
        _ptrs, weights = self._mlp_weight_ptrs()
        if torch.is_grad_enabled() and (h.requires_grad or any(t.requires_grad for t in (*gru, *weights))):
            cfg = {"H": self._hidden_dim, "n_hidden": int(self._num_edge_MLP_hidden_layers), "flags": self._flags(),
                   "agg": self._aggregation_fn.code, "path": _ffi.PATH[self._path]}
            return _GGNNFunction.apply(h, prepared, cfg, *gru, *weights)
        out = torch.empty((prepared.num_nodes, self._hidden_dim), dtype=torch.float32, device=h.device)
        ptrs, _keep = self._mlp_weight_ptrs()
        _ffi.check(_ffi.lib().tfgnn_b200_ggnn_fwd(
            prepared.handle, h.data_ptr(), int(h.shape[1]), ptrs, int(self._num_edge_MLP_hidden_layers),
            self._hidden_dim, self._flags(), self._aggregation_fn.code, self._gru_kernel.value.data_ptr(),
            self._gru_recurrent_kernel.value.data_ptr(), self._gru_bias.value.data_ptr(),
            _ffi.PATH[self._path], out.data_ptr(), stream_ptr()))
        return out

    def set_weights_from_oracle_dict(self, w: Dict[str, Any]) -> None:
        super().set_weights_from_oracle_dict(w)
        self._gru_kernel.assign(w["gru_kernel"])
        self._gru_recurrent_kernel.assign(w["gru_recurrent_kernel"])
        self._gru_bias.assign(w["gru_bias"])
