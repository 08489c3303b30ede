#!/bin/bash
# One gpurun call = one measurement session (GPU minutes are the scarce resource): parity suites in separate processes
# (a faulting kernel must not take the other suites down), the official bench line, per-workload lines, ncu launch lists.
# Usage under gpurun:  bash tools/gpu_session.sh <tag> [sections...]   sections: tests ab bench wl ncu full
set -u
TAG=${1:-r2}; shift || true
SECTIONS=${*:-"tests ab bench wl ncu"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
has() { [[ " $SECTIONS " == *" $1 "* ]]; }

if has tests; then
  for f in test_gpu_parity test_gpu_batch_builder test_gpu_graph_ops test_gpu_scale test_tf_golden; do
    timeout 900 python -m pytest tests/$f.py -m gpu -q --maxfail=8 --durations=8 -s > $OUT/pytest_$f.log 2>&1
    echo "== $f: rc=$? $(tail -1 $OUT/pytest_$f.log)"
    grep -E "^(FAILED|ERROR)|rel err|ms per batch" $OUT/pytest_$f.log | head -20
  done
fi
if has ab; then
  for v in 0 1; do
    TFGNN_B200_CORR_BF16=$v timeout 300 python bench.py --steps 20 --warmup 5 --skip-e2e --skip-cpu-baseline --skip-secondary \
      --no-clock-sampler > $OUT/ab_corr_bf16_$v.json 2> $OUT/ab_corr_bf16_$v.err
    echo "== corr_bf16=$v cfg2: $(grep -o '"ms_per_step": [0-9.]*' $OUT/ab_corr_bf16_$v.json | head -1)"
    TFGNN_B200_CORR_BF16=$v timeout 300 python bench.py --workload h320 --steps 20 --warmup 5 --skip-e2e --skip-cpu-baseline \
      --no-clock-sampler > $OUT/ab_h320_corr_bf16_$v.json 2> $OUT/ab_h320_corr_bf16_$v.err
    echo "== corr_bf16=$v h320: $(grep -o '"ms_per_step": [0-9.]*' $OUT/ab_h320_corr_bf16_$v.json | head -1)"
  done
fi
if has bench; then
  timeout 600 python bench.py > $OUT/bench_full.json 2> $OUT/bench_full.err
  echo "== bench rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_full.json").read().strip().splitlines()[-1])
    print("cfg2 ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e ms", d["e2e"]["ms_per_step"], "e2e val", d["e2e"]["value"])
    for k, v in (d.get("secondary") or {}).items():
        print(k, "ms", v["ms_per_layer"], "frac", v["roofline"]["frac"], "e2e ms", v.get("e2e", {}).get("ms_per_step"), "prep", v["prepare_ms"], v["prepare_wall_ms"])
    print("cpu", d["cpu_baseline"]["value"], "prep", d["config"]["prepare_ms"], d["config"]["prepare_wall_ms"], d["config"]["prepare_first_call_ms"])
except Exception as e:
    print("parse failed", e); print(open("$OUT/bench_full.err").read()[-2000:])
PY
fi
if has wl; then
  for wl in ${WLS:-cfg4 cfg3 cfg5_shard}; do
    timeout 400 python bench.py --workload $wl --skip-cpu-baseline --skip-e2e --no-clock-sampler --steps 10 > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err
    echo "== $wl: $(grep -o '"ms_per_step": [0-9.]*' $OUT/bench_$wl.json | head -1) $(grep -o '"frac": [0-9.]*' $OUT/bench_$wl.json | head -1)"
    tail -2 $OUT/bench_$wl.err
  done
fi
if has ncu; then
  for wl in ${NCU_WLS:-cfg1 cfg4 cfg5_shard}; do
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_$wl.csv \
      python bench.py --workload $wl --steps 3 --warmup 3 --skip-e2e --skip-cpu-baseline --no-clock-sampler > $OUT/launches_$wl.log 2>&1
    echo "== launches $wl: $(wc -l < $OUT/launches_$wl.csv) lines"
  done
fi
if has full; then
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:fused_rgcn -s 3 -c 1 -o $OUT/prof_fused -f \
    python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu-baseline --skip-secondary --no-clock-sampler > $OUT/ncu_full.log 2>&1
  ls -la $OUT/*.ncu-rep
fi
if has gemmprof; then
  timeout 300 python tools/bench_gemm.py --paths sorted_tc > $OUT/gemm_micro.txt 2>&1; cat $OUT/gemm_micro.txt
  if [ -n "${GEMM_NO_NCU:-}" ]; then echo "session $TAG done (gemm micro only)"; exit 0; fi
  for shp in 2000000,320,320 500000,128,384; do
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o $OUT/prof_gemm_${shp//,/_} -f \
      python tools/bench_gemm.py --shapes $shp --paths sorted_tc > $OUT/ncu_gemm_${shp//,/_}.log 2>&1
  done
  ls -la $OUT/*.ncu-rep
fi
echo "session $TAG done"
