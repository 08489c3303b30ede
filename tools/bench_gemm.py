"""Micro-benchmark of tfgnn_b200_dense_fwd (tcgen05 3xTF32 vs SIMT) for a few shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tf2_gnn_b200 import _ffi  # noqa: E402
from tf2_gnn_b200.runtime import stream_ptr  # noqa: E402

shapes = [(500_000, 128, 384), (500_000, 128, 128), (500_000, 640, 128), (1_000_000, 1024, 256),
          (1_000_000, 256, 1024), (2_000_000, 320, 320), (8000, 960, 320)]
paths = ("sorted_tc", "sorted")
if "--shapes" in sys.argv:      # e.g. --shapes 2000000,320,320;500000,128,384
    shapes = [tuple(int(x) for x in s.split(",")) for s in sys.argv[sys.argv.index("--shapes") + 1].split(";")]
if "--paths" in sys.argv:
    paths = tuple(sys.argv[sys.argv.index("--paths") + 1].split(","))
for V, K, N in shapes:
    x = torch.rand((V, K), device="cuda") - 0.5
    w = torch.rand((K, N), device="cuda") - 0.5
    out = torch.empty((V, N), device="cuda")
    for path in paths:
        def run():
            _ffi.check(_ffi.lib().tfgnn_b200_dense_fwd(x.data_ptr(), w.data_ptr(), out.data_ptr(), V, K, N, 1,
                                                        _ffi.PATH[path], stream_ptr()))
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"dense [{V}x{K}]x[{K}x{N}] {path:10s} {ms:8.3f} ms  {2 * V * K * N / ms / 1e9:8.1f} TFLOP/s(fp32-equiv)  "
              f"{(V * K + V * N) * 4 / ms / 1e6:7.0f} GB/s", flush=True)
