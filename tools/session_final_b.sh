# ncu evidence of a build: one --set full capture of the dominant kernel per bench workload + launch lists of the variants.
set -u
OUT=gpurun_out/${1:-final_b}; mkdir -p $OUT
B="python bench.py --skip-e2e --skip-cpu-baseline --skip-secondary --no-clock-sampler"
for wl in cfg2 h320 cfg1; do
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:fused_rgcn -s 3 -c 1 -o $OUT/prof_fused_$wl -f \
    $B --workload $wl --steps 3 --warmup 3 > $OUT/ncu_full_$wl.log 2>&1
done
timeout 500 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o $OUT/prof_gemm_gru -f \
  $B --workload cfg4 --steps 3 --warmup 3 > $OUT/ncu_full_gru.log 2>&1
for wl in cfg2 cfg1 cfg3 cfg4 cfg5_shard; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file $OUT/launches_$wl.csv \
    $B --workload $wl --steps 3 --warmup 3 > $OUT/launches_$wl.log 2>&1
  echo "== launches $wl: $(wc -l < $OUT/launches_$wl.csv) lines"
done
ls -la $OUT/*.ncu-rep
echo "session final_b done"
