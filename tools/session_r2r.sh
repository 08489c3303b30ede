set -u
OUT=gpurun_out/r2r; mkdir -p $OUT
B="python bench.py --skip-e2e --skip-cpu-baseline --skip-secondary --no-clock-sampler"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "ggnn or stack" > $OUT/pytest_ggnn.log 2>&1; echo "== ggnn parity: $(tail -1 $OUT/pytest_ggnn.log)"; grep -E "^(FAILED|ERROR)|Error" $OUT/pytest_ggnn.log | head
timeout 600 python -m pytest tests/test_gpu_graph_ops.py -m gpu -q > $OUT/pytest_ops.log 2>&1; echo "== graph ops: $(tail -1 $OUT/pytest_ops.log)"; grep -E "^(FAILED|ERROR)" $OUT/pytest_ops.log | head
timeout 300 python -m pytest tests/test_gpu_scale.py -m gpu -q -s -k "cfg4" > $OUT/pytest_scale.log 2>&1; echo "== scale cfg4: $(tail -1 $OUT/pytest_scale.log)"; grep "rel err" $OUT/pytest_scale.log
for v in 1 0 1; do
  TFGNN_B200_GGNN_FUSED_GRU=$v timeout 300 $B --workload cfg4 --steps 10 > $OUT/bench_cfg4_$v.json 2> $OUT/bench_cfg4_$v.err
  echo "== cfg4 fused_gru=$v: $(grep -o '"ms_per_step": [0-9.]*' $OUT/bench_cfg4_$v.json | head -1) $(grep -o '"frac": [0-9.]*' $OUT/bench_cfg4_$v.json | head -1) $(tail -1 $OUT/bench_cfg4_$v.err | cut -c1-200)"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_cfg4.csv $B --workload cfg4 --steps 3 --warmup 3 > $OUT/launches_cfg4.log 2>&1
python tools/launch_summary.py $OUT/launches_cfg4.csv 2>/dev/null | tail -8
echo "session r2r done"
