set -u
OUT=gpurun_out/r2o; mkdir -p $OUT
B="python bench.py --skip-e2e --skip-cpu-baseline --skip-secondary --no-clock-sampler"
for f in test_gpu_parity test_gpu_graph_ops; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --maxfail=8 > $OUT/pytest_$f.log 2>&1; echo "== $f: $(tail -1 $OUT/pytest_$f.log)"
  grep -E "^(FAILED|ERROR)" $OUT/pytest_$f.log | head -12
done
for wl in cfg2 h320 cfg1; do
  for v in "1 1" "1 0" "0 0"; do
    set -- $v
    TFGNN_B200_EPI_DIRECT=$1 TFGNN_B200_EPI_HELPERS=$2 timeout 200 $B --workload $wl --steps 15 --warmup 4 > $OUT/epi_${wl}_$1$2.json 2> $OUT/epi_${wl}_$1$2.err
    echo "== $wl direct=$1 helpers=$2: $(grep -o '"ms_per_step": [0-9.]*' $OUT/epi_${wl}_$1$2.json | head -1)"
  done
done
for wl in cfg4 cfg3 cfg5_shard; do
  timeout 400 $B --workload $wl --steps 10 > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err
  echo "== $wl: $(grep -o '"ms_per_step": [0-9.]*' $OUT/bench_$wl.json | head -1) $(grep -o '"frac": [0-9.]*' $OUT/bench_$wl.json | head -1)"
done
for wl in cfg2 h320 cfg1; do
  TFGNN_B200_FUSED_TRACE=$OUT/trace_$wl.bin timeout 200 $B --workload $wl --steps 1 --warmup 1 > $OUT/trace_$wl.json 2> $OUT/trace_$wl.err
done
timeout 300 python -m pytest tests/test_gpu_scale.py -m gpu -q -s > $OUT/pytest_scale.log 2>&1; echo "== scale: $(tail -1 $OUT/pytest_scale.log)"; grep "rel err" $OUT/pytest_scale.log
echo "session r2o done"
