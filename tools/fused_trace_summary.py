"""Summarise a TFGNN_B200_FUSED_TRACE dump (debug timeline of the fused layer kernel, fused_rgcn.cu: fu_trace()).

  TFGNN_B200_FUSED_TRACE=trace.bin python bench.py --workload cfg1 --steps 1 --warmup 1 --skip-e2e ...
  python tools/fused_trace_summary.py trace.bin [sm_mhz]

Stamps are SM-local clock64 values: every column is reported relative to the CTA's own kernel entry, in microseconds,
as the median / min / max over CTAs."""
import sys

import numpy as np


def main():
    path = sys.argv[1]
    mhz = float(sys.argv[2]) if len(sys.argv) > 2 else 1965.0
    raw = np.fromfile(path, dtype=np.int64)
    grid, slots, split, ctas = (int(x) for x in raw[:4])
    t = raw[4:].reshape(grid, slots).astype(np.float64)
    entry = t[:, 0:1]
    rel = np.where(t > 0, (t - entry) / mhz, np.nan)
    gt = t[:, 3]
    print(f"grid {grid}, split {split}, ctas/cluster {ctas}; kernel entry skew across CTAs (globaltimer) "
          f"{(gt.max() - gt.min()) / 1e3:.1f} us")
    if np.all(t[:, 4] > 0):
        dur_ns = t[:, 4].max() - gt.min()
        cyc = np.median(t[:, 2] - t[:, 0])
        print(f"kernel duration (globaltimer, first entry -> last exit) {dur_ns / 1e3:.1f} us; median CTA {cyc:.0f} cycles "
              f"-> SM clock ~{cyc / np.median(t[:, 4] - gt) * 1e3:.0f} MHz (pass it as argv[2])")

    def line(name, idx):
        col = rel[:, idx]
        if np.all(np.isnan(col)):
            return
        print(f"{name:28s} median {np.nanmedian(col):9.2f}  min {np.nanmin(col):9.2f}  max {np.nanmax(col):9.2f}  "
              f"(n={int(np.sum(~np.isnan(col)))})")

    line("set-up done", 1)
    for cc in range(40):
        line(f"gather warp0 call {cc} done", 8 + cc)
    for u in range(16):
        line(f"TMA got first slot, unit {u}", 112 + u)
    for k in range(16):
        line(f"MMA tile {k} first issue", 48 + 2 * k)
        line(f"MMA tile {k} last commit", 49 + 2 * k)
        line(f"epilogue tile {k} start", 80 + 2 * k)
        line(f"epilogue tile {k} end", 81 + 2 * k)
    line("exit", 2)


if __name__ == "__main__":
    main()
