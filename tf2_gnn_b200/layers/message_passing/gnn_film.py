"""GNN_FiLM — mirror of tf2_gnn/layers/message_passing/gnn_film.py:13-108 on the B200 path."""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch

from ... import _ffi
from ...runtime import PreparedBatch, stream_ptr
from .gnn_edge_mlp import EdgeMLP, GNN_Edge_MLP
from ..differentiable import edge_mlp_family_forward
from ..node_ops import _needs_grad
from .message_passing import MessagePassingInput, _last_dim, register_message_passing_implementation


@register_message_passing_implementation
class GNN_FiLM(GNN_Edge_MLP):
    """h'_v = sum_l sum_{(u,v) in A_l} sigma(1/c_{v,l} * gamma_{l,v} * (W_l h_u) + beta_{l,v}),
    [gamma|beta] = F_l(h_v)  (gnn_film.py:14-47)."""

    @classmethod
    def get_default_hyperparameters(cls):
        these_hypers = {
            "use_target_state_as_input": False,
            "normalize_by_num_incoming": False,
            "num_edge_MLP_hidden_layers": 0,
            "film_parameter_MLP_hidden_layers": [],
        }
        mp_hypers = super().get_default_hyperparameters()
        mp_hypers.update(these_hypers)
        return mp_hypers

    def __init__(self, params: Dict[str, Any], **kwargs):
        super().__init__(params, **kwargs)
        self._film_parameter_MLP_hidden_layers = params["film_parameter_MLP_hidden_layers"]
        self._edge_type_film_layer_computations: List[EdgeMLP] = []

    def build(self, input_shapes: MessagePassingInput):
        D = _last_dim(input_shapes.node_embeddings)
        for i in range(len(input_shapes.adjacency_lists)):
            self._edge_type_film_layer_computations.append(
                EdgeMLP(self, f"edge_type_{i}-FiLM", D, 2 * self._hidden_dim,
                        list(self._film_parameter_MLP_hidden_layers)))
        super().build(input_shapes)

    def call(self, inputs: MessagePassingInput, training: bool = False,
             prepared: Optional[PreparedBatch] = None):
        h, prepared = self._device_inputs(inputs, prepared)
        self._check_types(prepared)
        if _needs_grad(h, *[v.value for v in self.variables]):
            # training: the reference's literal op order with per-op backward kernels (layers/differentiable.py)
            return edge_mlp_family_forward(
                self, h, prepared,
                film_kernels=[[v.value for v in m.layers] for m in self._edge_type_film_layer_computations])
        if any(m.num_hidden_layers for m in self._edge_type_film_layer_computations):
            raise NotImplementedError("film_parameter_MLP_hidden_layers != [] is not built yet")
        out = torch.empty((prepared.num_nodes, self._hidden_dim), dtype=torch.float32, device=h.device)
        ptrs, _keep = self._mlp_weight_ptrs()
        film = [m.layers[0].value for m in self._edge_type_film_layer_computations]
        _ffi.check(_ffi.lib().tfgnn_b200_film_fwd(
            prepared.handle, h.data_ptr(), int(h.shape[1]), ptrs, int(self._num_edge_MLP_hidden_layers),
            _ffi.ptr_array(film), self._hidden_dim, self._flags(), self._aggregation_fn.code,
            self._activation_fn.code, _ffi.PATH[self._path], out.data_ptr(), stream_ptr()))
        return out

    def set_weights_from_oracle_dict(self, w: Dict[str, Any]) -> None:
        super().set_weights_from_oracle_dict(w)
        for mlp, mats in zip(self._edge_type_film_layer_computations, w["film_mlps"]):
            for var, m in zip(mlp.layers, mats):
                var.assign(m)
