set -u
OUT=gpurun_out/r2n; mkdir -p $OUT
B="python bench.py --skip-e2e --skip-cpu-baseline --skip-secondary --no-clock-sampler"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "gnn_stack_parity" > $OUT/pytest_stack.log 2>&1; echo "== stack parity: $(tail -1 $OUT/pytest_stack.log)"
for wl in cfg1 cfg2 h320; do
  TFGNN_B200_FUSED_TRACE=$OUT/trace_$wl.bin timeout 200 $B --workload $wl --steps 1 --warmup 1 > $OUT/trace_$wl.json 2> $OUT/trace_$wl.err; echo "== trace $wl rc=$? $(ls -la $OUT/trace_$wl.bin | awk '{print $5}')"
done
for wl in cfg2 h320 cfg1; do
  for sk in 0 8 16 24 2 4 1; do
    TFGNN_B200_DEBUG_SKIP=$sk timeout 200 $B --workload $wl --steps 10 --warmup 3 > $OUT/skip_${wl}_$sk.json 2> $OUT/skip_${wl}_$sk.err
    echo "== $wl skip=$sk: $(grep -o '"ms_per_step": [0-9.]*' $OUT/skip_${wl}_$sk.json | head -1)"
  done
done
for wl in h320 cfg1; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_rgcn -s 3 -c 1 -o $OUT/prof_fused_$wl -f \
    $B --workload $wl --steps 3 --warmup 3 > $OUT/ncu_full_$wl.log 2>&1
done
ls -la $OUT/*.ncu-rep
echo "session r2n done"
