"""Multi-GPU check + timing of the target-range sharded path (SURVEY.md §8e case 2) over NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tools/run_sharded_nccl.py [--nodes 1000000 --edges-per-type 5000000 --types 4 --hidden 256 --layers 2]

Every rank builds the same seeded graph, owns one target range (tfgnn_b200_prepare_sharded on the FULL edge
lists), and runs `layers` RGCN layers with ONE all-gather of the node-state shards per layer
(torch.distributed.all_gather_into_tensor, NCCL over NVLink).  Rank 0 also runs the unsharded layers and
checks the gathered result against it, then prints one JSON line with per-layer times (max over ranks).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tf2_gnn_b200 import sharding  # noqa: E402
from tf2_gnn_b200.layers import MessagePassingInput, RGCN  # noqa: E402
from tf2_gnn_b200.runtime import PreparedBatch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=400_000)
    ap.add_argument("--edges-per-type", type=int, default=2_000_000)
    ap.add_argument("--types", type=int, default=4)
    ap.add_argument("--hidden", type=int, default=256)
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    V, L, H = args.nodes, args.types, args.hidden
    rng = np.random.default_rng(7)
    adjs = [rng.integers(0, V, size=(args.edges_per_type, 2), dtype=np.int32) for _ in range(L)]
    h0 = (rng.random((V, H), dtype=np.float32) * 2 - 1)
    lim = np.sqrt(6.0 / (2 * H))
    ws = [[((rng.random((H, H), dtype=np.float32) * 2 - 1) * lim) for _ in range(L)] for _ in range(args.layers)]
    deg = sum(np.bincount(a[:, 1], minlength=V) for a in adjs)
    bounds = sharding.partition_target_range(V, world, deg)
    lo, hi = bounds[rank]
    adj_dev = tuple(torch.from_numpy(a).cuda() for a in adjs)
    params = RGCN.get_default_hyperparameters()
    params["hidden_dim"] = H
    layers = []
    for w in ws:
        layer = RGCN(params)
        layer.build(MessagePassingInput((None, H), tuple((None, 2) for _ in range(L))))
        layer.set_weights_from_oracle_dict({"edge_mlps": [[x] for x in w]})
        layers.append(layer)
    shard = PreparedBatch(adj_dev, V, target_range=(lo, hi))
    h_local0 = torch.from_numpy(h0[lo:hi]).cuda()

    def forward_sharded():
        h_local = h_local0
        for layer in layers:
            h_full = sharding.all_gather_node_states(h_local, bounds)          # one collective per layer
            h_local = layer(MessagePassingInput(h_full, adj_dev), prepared=shard)
        return h_local

    out_local = forward_sharded()
    torch.cuda.synchronize()
    dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        out_local = forward_sharded()
    ev1.record()
    torch.cuda.synchronize()
    t = torch.tensor([ev0.elapsed_time(ev1) / args.steps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = sharding.all_gather_node_states(out_local, bounds)
    ok, err = True, 0.0
    if rank == 0:
        full = PreparedBatch(adj_dev, V)
        h = torch.from_numpy(h0).cuda()
        for layer in layers:
            h = layer(MessagePassingInput(h, adj_dev), prepared=full)
        err = float((gathered - h).abs().max() / h.abs().max())
        ok = err <= 2e-6
        M = L * args.edges_per_type
        print(json.dumps({"check": "target-range sharded RGCN == unsharded", "world_size": world, "nodes": V,
                          "edges": M, "hidden": H, "layers": args.layers, "max_rel_err": err, "ok": ok,
                          "ms_per_forward_max_over_ranks": float(t.item()),
                          "edges_per_s": M * args.layers / (float(t.item()) * 1e-3),
                          "allgather_bytes_per_layer_per_rank": int(V * H * 4)}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
