"""tf2_gnn_b200 — B200-native (sm_100a) implementation of the tf2_gnn message-passing hot path,
behind the reference's layer API.  See DESIGN.md."""
__version__ = "0.1.0"
