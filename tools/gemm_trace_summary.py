"""Summarise a TFGNN_B200_GEMM_TRACE dump (debug timeline of gemm_tc_kernel, gemm_tc.cu: tc_trace()).

  TFGNN_B200_GEMM_TRACE=trace.bin python bench.py --workload cfg4 --steps 1 --warmup 1 --skip-e2e ...
  python tools/gemm_trace_summary.py trace.bin [sm_mhz]

Per K block of the first 60 of a CTA, medians over CTAs, in microseconds since kernel entry (SM-local clock64)."""
import sys

import numpy as np


def main():
    mhz = float(sys.argv[2]) if len(sys.argv) > 2 else 1965.0
    raw = np.fromfile(sys.argv[1], dtype=np.int64)
    grid, slots, nkb, stages = (int(x) for x in raw[:4])
    t = raw[4:].reshape(grid, slots).astype(np.float64)
    rel = np.where(t > 0, (t - t[:, 240:241]) / mhz, np.nan)
    med = np.nanmedian(rel, axis=0)
    print(f"grid {grid}, {nkb} K blocks per tile, {stages} stages; exit {med[243]:.1f} us; first tile epilogue {med[241]:.2f} -> {med[242]:.2f}")
    print("  kb  tma_issue  bytes_seen  mma_issue  | load  split  ||  stage_free(for kb)  d(mma_issue)")
    prev = np.nan
    for i in range(60):
        a, b, c, d = med[4 * i], med[4 * i + 1], med[4 * i + 2], med[4 * i + 3]
        if np.isnan(a):
            break
        print(f"{i:4d} {a:9.2f} {b:10.2f} {c:10.2f}  | {b - a:5.2f} {c - b:5.2f}  || {d:9.2f}  {c - prev:6.2f}")
        prev = c


if __name__ == "__main__":
    main()
