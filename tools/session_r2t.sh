set -u
OUT=gpurun_out/r2t; mkdir -p $OUT
B="python bench.py --skip-e2e --skip-cpu-baseline --skip-secondary --no-clock-sampler"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "rgat or stack or film or variants or activations" > $OUT/pytest_rgat.log 2>&1; echo "== rgat parity: $(tail -1 $OUT/pytest_rgat.log)"; grep -E "^(FAILED|ERROR)|Error" $OUT/pytest_rgat.log | head
timeout 300 python -m pytest tests/test_gpu_scale.py -m gpu -q -s -k "cfg3" > $OUT/pytest_scale.log 2>&1; echo "== scale cfg3: $(tail -1 $OUT/pytest_scale.log)"; grep "rel err" $OUT/pytest_scale.log
timeout 300 python -m pytest tests/test_gpu_graph_ops.py -m gpu -q -k "rgat" > $OUT/pytest_ops.log 2>&1; echo "== graph ops rgat: $(tail -1 $OUT/pytest_ops.log)"
for v in 1 0 1; do
  TFGNN_B200_RGAT_FUSED_SCORES=$v timeout 300 $B --workload cfg3 --steps 8 > $OUT/bench_cfg3_$v.json 2> $OUT/bench_cfg3_$v.err
  echo "== cfg3 fused_scores=$v: $(grep -o '"ms_per_step": [0-9.]*' $OUT/bench_cfg3_$v.json | head -1) $(grep -o '"frac": [0-9.]*' $OUT/bench_cfg3_$v.json | head -1) $(tail -1 $OUT/bench_cfg3_$v.err | cut -c1-200)"
done
echo "session r2t done"
