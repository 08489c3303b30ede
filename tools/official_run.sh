# Round-end measurement set: GPU parity suite, the official bench lines, the ncu launch list.  Run under gpurun.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_full.log 2>&1; tail -c 600 gpurun_out/bench_full.log; echo
for wl in h320 cfg1 cfg4 cfg3; do
  timeout 300 python bench.py --workload $wl --skip-cpu-baseline > gpurun_out/bench_$wl.log 2>&1
  grep -o "\"ms_per_step\": [0-9.]*" gpurun_out/bench_$wl.log | head -1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 5 --warmup 3 --skip-e2e --skip-cpu-baseline --no-clock-sampler > gpurun_out/launches_final.log 2>&1
tail -2 gpurun_out/launches_final.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_rgcn -s 3 -c 1 -o gpurun_out/prof_fused_final2 -f python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu-baseline --no-clock-sampler > gpurun_out/ncu_final2.log 2>&1
ls -la gpurun_out/prof_fused_final2.ncu-rep
