"""The debug-timeline readers (tools/fused_trace_summary.py, tools/gemm_trace_summary.py) on synthetic dumps in the formats
fused_rgcn.cu (fu_trace) and gemm_tc.cu (tc_trace) write: a 4-word int64 header, then grid x slots stamps."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(tool, path, *args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), path, *args], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return r.stdout


def test_fused_trace_summary_reads_a_dump(tmp_path):
    grid, slots = 4, 128
    t = np.zeros((grid, slots), np.int64)
    for c in range(grid):
        base = 1_000_000 * (c + 1)
        t[c, 0] = base                       # entry
        t[c, 1] = base + 2000                # set-up done
        t[c, 2] = base + 2_000_000           # exit
        t[c, 3] = 5_000_000_000 + c          # globaltimer at entry (ns)
        t[c, 4] = 5_000_000_000 + 1_000_000  # globaltimer at exit: 1 ms
        t[c, 8] = base + 20_000              # gather call 0
        t[c, 112] = base + 22_000            # TMA got the first slot of unit 0
        if c % 2 == 0:                       # leader CTAs only
            t[c, 48], t[c, 49] = base + 25_000, base + 90_000
        t[c, 80], t[c, 81] = base + 91_000, base + 100_000
    p = tmp_path / "fused.bin"
    np.concatenate([np.array([grid, slots, 0, 2], np.int64), t.ravel()]).tofile(p)
    out = _run("fused_trace_summary.py", str(p), "2000")
    assert "grid 4" in out and "ctas/cluster 2" in out
    assert "SM clock ~2000 MHz" in out       # 2.0 M cycles in 1.0 ms
    assert "gather warp0 call 0 done" in out and "(n=2)" in out   # the MMA lines count the two leaders only
    line = [l for l in out.splitlines() if l.startswith("set-up done")][0]
    assert abs(float(line.split()[3]) - 1.0) < 1e-6               # 2000 cycles at 2000 MHz = 1 us


def test_gemm_trace_summary_reads_a_dump(tmp_path):
    grid, slots, nkb, stages = 3, 256, 10, 4
    t = np.zeros((grid, slots), np.int64)
    for c in range(grid):
        base = 7_000 * (c + 1)
        t[c, 240] = base
        for i in range(12):
            t[c, 4 * i] = base + 1000 * i + 100          # TMA issue
            t[c, 4 * i + 1] = base + 1000 * i + 800      # bytes seen
            t[c, 4 * i + 2] = base + 1000 * i + 1500     # MMA issue
            t[c, 4 * i + 3] = base + 1000 * i + 90       # stage free
        t[c, 241], t[c, 242], t[c, 243] = base + 11_000, base + 14_000, base + 400_000
    p = tmp_path / "gemm.bin"
    np.concatenate([np.array([grid, slots, nkb, stages], np.int64), t.ravel()]).tofile(p)
    out = _run("gemm_trace_summary.py", str(p), "1000")
    assert "10 K blocks per tile, 4 stages" in out
    rows = [l.split() for l in out.splitlines() if l.strip() and l.split()[0].isdigit()]
    assert len(rows) == 12
    assert abs(float(rows[3][1]) - 3.1) < 1e-6 and abs(float(rows[3][3]) - 4.5) < 1e-6   # us at 1000 MHz
    assert abs(float(rows[5][-1]) - 1.0) < 1e-6                                            # MMA issue cadence
