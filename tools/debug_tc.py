"""Diagnostics for the tcgen05 GEMM (run on the GPU box): structured inputs that reveal operand
layout / descriptor mistakes in one round trip."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tf2_gnn_b200 import _ffi  # noqa: E402
from tf2_gnn_b200.runtime import stream_ptr  # noqa: E402


def dense(x, w, path):
    V, K = x.shape
    N = w.shape[1]
    xt, wt = torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda()
    out = torch.full((V, N), float("nan"), dtype=torch.float32, device="cuda")
    _ffi.check(_ffi.lib().tfgnn_b200_dense_fwd(xt.data_ptr(), wt.data_ptr(), out.data_ptr(), V, K, N, 0,
                                                _ffi.PATH[path], stream_ptr()))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def main():
    rng = np.random.default_rng(0)
    for (V, K, N) in [(128, 32, 64), (128, 64, 64), (256, 32, 256), (300, 96, 320), (1000, 1024, 256)]:
        x = rng.uniform(-1, 1, (V, K)).astype(np.float32)
        w = rng.uniform(-1, 1, (K, N)).astype(np.float32)
        ref = x.astype(np.float64) @ w.astype(np.float64)
        got = dense(x, w, "sorted_tc")
        err = np.abs(got - ref)
        print(f"[{V}x{K}x{N}] max|err|={np.nanmax(err):.3e} rel={np.nanmax(err) / np.abs(ref).max():.3e} "
              f"nan={int(np.isnan(got).sum())}")
        if not (np.nanmax(err) / np.abs(ref).max() < 1e-5) or np.isnan(got).any():
            bad = (err > 1e-4 * np.abs(ref).max()) | np.isnan(got)
            rows = np.where(bad.any(axis=1))[0]
            cols = np.where(bad.any(axis=0))[0]
            print("   bad rows:", rows[:16], "... count", len(rows), " bad cols:", cols[:16], "... count", len(cols))
            # K mapping probe: x = e_{k0} for all rows -> out rows should equal w[k0]
            for k0 in (0, 1, 7, 8, 9, 31, K - 1):
                xe = np.zeros((V, K), np.float32)
                xe[:, k0] = 1.0
                g = dense(xe, w, "sorted_tc")
                match = [kk for kk in range(K) if np.allclose(g[0], w[kk], atol=1e-3)]
                print(f"   k-probe k0={k0}: row0 matches w[{match}]  row0[:4]={g[0][:4]} want={w[k0][:4]}")
            # M mapping probe: x[m, 0] = m -> out[m, n] = m * w[0, n]
            xe = np.zeros((V, K), np.float32)
            xe[:, 0] = np.arange(V)
            g = dense(xe, w, "sorted_tc")
            rows_got = np.round(g[:, 0] / w[0, 0]).astype(np.int64)
            print("   m-probe: recovered row ids:", rows_got[:16], "...", rows_got[-4:])
            break


if __name__ == "__main__":
    main()
