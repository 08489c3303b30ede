"""RGAT — mirror of tf2_gnn/layers/message_passing/rgat.py:11-163 on the B200 path."""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch

from ... import _ffi
from ...runtime import PreparedBatch, stream_ptr
from ..node_ops import _needs_grad
from .message_passing import (MessagePassing, MessagePassingInput, Variable, _last_dim,
                              register_message_passing_implementation)


@register_message_passing_implementation
class RGAT(MessagePassing):
    """Relational graph attention (rgat.py:12-51): per type a bias-free Dense W_l [D,H] applied to
    source and target states, K-head scores leaky_relu(a_l . [W_l h_u || W_l h_v]), softmax over ALL
    incoming edges of a node (all types jointly, rgat.py:135-151), weighted sum, activation."""

    @classmethod
    def get_default_hyperparameters(cls):
        these_hypers = {"num_heads": 3}
        mp_hypers = super().get_default_hyperparameters()
        mp_hypers.update(these_hypers)
        return mp_hypers

    def __init__(self, params: Dict[str, Any], **kwargs):
        super().__init__(params, **kwargs)
        self._num_heads: int = params["num_heads"]
        self._edge_type_to_message_computation_layer: List[Variable] = []
        self._edge_type_to_attention_parameters: List[Variable] = []

    def build(self, input_shapes: MessagePassingInput):
        D = _last_dim(input_shapes.node_embeddings)
        per_head_dim = self._hidden_dim // self._num_heads
        for i in range(len(input_shapes.adjacency_lists)):
            self._edge_type_to_message_computation_layer.append(
                self.add_weight(f"edge_type_{i}/Edge_weight_{i}/kernel:0", (D, self._hidden_dim)))
            self._edge_type_to_attention_parameters.append(
                self.add_weight(f"edge_type_{i}/Edge_attention_parameters_{i}:0",
                                (self._num_heads, 2 * per_head_dim)))
        super().build(input_shapes)

    def call(self, inputs: MessagePassingInput, training: bool = False,
             prepared: Optional[PreparedBatch] = None):
        h, prepared = self._device_inputs(inputs, prepared)
        if _needs_grad(h, *[v.value for v in self.variables]):
            # training: the reference's literal op order with per-op backward kernels (layers/differentiable.py)
            from ..differentiable import rgat_forward
            return rgat_forward(self, h, prepared)
        if prepared.num_edge_types != len(self._edge_type_to_message_computation_layer):
            raise ValueError("number of adjacency lists differs from the number the layer was built for")
        if self._hidden_dim % self._num_heads:
            raise ValueError("hidden_dim must be divisible by num_heads (rgat.py:72)")
        out = torch.empty((prepared.num_nodes, self._hidden_dim), dtype=torch.float32, device=h.device)
        kernels = [v.value for v in self._edge_type_to_message_computation_layer]
        att = [v.value for v in self._edge_type_to_attention_parameters]
        _ffi.check(_ffi.lib().tfgnn_b200_rgat_fwd(
            prepared.handle, h.data_ptr(), int(h.shape[1]), _ffi.ptr_array(kernels), _ffi.ptr_array(att),
            self._hidden_dim, int(self._num_heads), self._activation_fn.code, _ffi.PATH[self._path],
            out.data_ptr(), stream_ptr()))
        return out

    def _message_function(self, *args, **kwargs):
        raise NotImplementedError("built-in layers run fused; _message_function is only a plugin hook")

    def set_weights_from_oracle_dict(self, w: Dict[str, Any]) -> None:
        for var, m in zip(self._edge_type_to_message_computation_layer, w["edge_kernels"]):
            var.assign(m)
        for var, m in zip(self._edge_type_to_attention_parameters, w["edge_attention"]):
            var.assign(m)
