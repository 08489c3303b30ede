"""Measurement of the backward pass at bench.py's workloads: forward and backward device time of one
RGCN (or GGNN) layer through the autograd hook, CUDA events, inputs resident in HBM.
  python tools/bench_backward.py [--workload cfg2] [--steps 10] [--warmup 3]
Algorithmic bytes of the backward: gather of h rows for A (recomputed) + scatter-side gather of dA rows + dOut/out reads +
dh write + weights, i.e. about twice the forward's."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tf2_gnn_b200.layers import MessagePassingInput, get_message_passing_class  # noqa: E402
from tf2_gnn_b200.runtime import PreparedBatch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    wl = bench.WORKLOADS[args.workload]
    V, H, L = wl["V"], wl["H"], len(wl["E"])
    h_np, adjs_np, w_np = bench.make_inputs(wl, seed=0)
    kind = wl["kind"]
    cls = get_message_passing_class(kind)
    params = cls.get_default_hyperparameters()
    params.update(wl.get("params", {}))
    params.update(hidden_dim=H)
    layer = cls(params)
    torch.manual_seed(1)
    layer.build(MessagePassingInput((None, H), tuple((None, 2) for _ in range(L))))
    for v in layer.variables:
        v.requires_grad_()
    dev = torch.device("cuda", 0)
    h = torch.from_numpy(h_np).to(dev).requires_grad_()
    adj = tuple(torch.from_numpy(a).to(dev) for a in adjs_np)
    prepared = PreparedBatch(adj, V)
    prepared.transposed()
    g = torch.rand((V, H), device=dev) * 2 - 1
    inp = MessagePassingInput(h, adj)
    fwd_ms, bwd_ms = [], []
    for i in range(args.warmup + args.steps):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        h.grad = None
        for v in layer.variables:
            v.value.grad = None
        e0.record()
        out = layer(inp, prepared=prepared)
        e1.record()
        out.backward(g)
        e2.record()
        torch.cuda.synchronize()
        if i >= args.warmup:
            fwd_ms.append(e0.elapsed_time(e1))
            bwd_ms.append(e1.elapsed_time(e2))
    M = sum(wl["E"])
    alg_fwd = bench.algorithmic_bytes(kind, V, wl["E"], H, H, params)
    print(json.dumps({
        "workload": wl["desc"], "kind": kind, "forward_ms": float(np.median(fwd_ms)), "backward_ms": float(np.median(bwd_ms)),
        "edges_per_s_fwd_bwd": M / ((np.median(fwd_ms) + np.median(bwd_ms)) * 1e-3),
        "forward_algorithmic_bytes": alg_fwd,
        "note": "backward = recompute A (CSR reduce) + TN GEMM dW (fp32 FFMA) + tensor-core GEMM dA + source-keyed CSR "
                "reduce dh; autograd hook overhead included; not tuned (two-kernel form, SIMT dW)"}), flush=True)


if __name__ == "__main__":
    main()
