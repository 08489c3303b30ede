"""RGCN — mirror of tf2_gnn/layers/message_passing/rgcn.py:12-62 on the B200 path."""
from __future__ import annotations

from typing import Any, Dict

from .gnn_edge_mlp import GNN_Edge_MLP
from .message_passing import register_message_passing_implementation


@register_message_passing_implementation
class RGCN(GNN_Edge_MLP):
    """h'_v = sigma(sum_l sum_{(u,v) in A_l} 1/c_{v,l} * (W_l h_u))  (rgcn.py:13-48), no basis
    decomposition: one dense W_l [D,H] per edge type, no bias (test_RGCN.py:35-39)."""

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
